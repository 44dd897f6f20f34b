// cc_emu.cpp -- TEST INFRASTRUCTURE: compiles the component kernels of metafast_b200/csrc/components.cuh for the HOST
// (g++; one emulated CUDA thread per host thread, grid of 1..N one-thread blocks, real atomics) and drives them with the same
// level loop as mfkc_kset_components_begin (metafast_b200/csrc/kset_api.inl), so that the kernel logic and the host
// grouping are checked against the oracle on machines without a GPU (tests/test_host.py).  It proves the arithmetic and
// the level logic, not the concurrency; the GPU tests cover the real launch.  Never part of the product.
#include <cuda_runtime.h>      // host_defines.h: __device__, __global__, __launch_bounds__, __align__ as (ignored) attributes
#include <stdint.h>
#include <string.h>

#include <vector>

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#include <thread>
// one emulated CUDA thread per host thread: block t of a grid of g_grid one-thread blocks.  With g_grid > 1 the kernels run
// truly concurrently (real atomics), which exercises the lock-free union-find under contention on the CPU's memory model.
static thread_local uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
static const dim3 blockDim(1, 1, 1);
static dim3 gridDim(1, 1, 1);
template <class T, class U, class V> static T atomicCAS(T *p, U cmp, V val) {
    T expected = (T)cmp;
    __atomic_compare_exchange_n(p, &expected, (T)val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;
}
template <class T, class U> static T atomicAdd(T *p, U v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_SEQ_CST); }
template <class K, class... A> static void launch(int grid, K kernel, A... args) {
    gridDim.x = (unsigned)grid;
    if (grid == 1) { blockIdx.x = 0; kernel(args...); return; }
    std::vector<std::thread> ts;
    for (int t = 0; t < grid; t++) ts.emplace_back([=] { blockIdx.x = (unsigned)t; kernel(args...); });
    for (auto &t : ts) t.join();
}

#include "../../metafast_b200/csrc/components.cuh"
#include "../../metafast_b200/csrc/components_host.h"

using namespace mfkc;

extern "C" int cc_emu_components(const unsigned long long *keys, const uint32_t *vals, uint64_t n, int k, long long b1, long long b2,
                                 uint64_t *n_comp, uint64_t *n_keys, int *levels, void **result, int grid) {
    if (grid < 1) grid = 1;
    std::vector<uint8_t> active(n ? n : 1);
    std::vector<uint32_t> label(n ? n : 1), thr_of(n ? n : 1), parent(n ? n : 1), size(n ? n : 1);
    const uint64_t cap = n * 2 + 64;
    std::vector<Slot> tab(cap);
    memset(tab.data(), 0xFF, cap * sizeof(Slot));
    CcIndex ix; ix.tab = tab.data(); ix.cap = cap;
    int lv = 0;
    if (n) {
        launch(grid, cc_index_build_kernel, keys, n, tab.data(), cap);
        launch(grid, cc_begin_kernel, vals, n, active.data(), label.data(), thr_of.data());
        for (int thr = 1; thr <= 32767; thr++) {
            unsigned long long counters[2] = {0, 0};
            launch(grid, cc_level_init_kernel, n, parent.data(), size.data());
            launch(grid, cc_union_kernel, keys, n, (const uint8_t *)active.data(), ix, k, parent.data());
            launch(grid, cc_count_kernel, n, (const uint8_t *)active.data(), parent.data(), size.data());
            launch(grid, cc_classify_kernel, vals, n, active.data(), (const uint32_t *)parent.data(), (const uint32_t *)size.data(), b1, b2, thr, label.data(), thr_of.data(), (unsigned long long *)counters);
            lv = thr;
            if (!counters[0]) break;
        }
    }
    CcResult *res = new CcResult();
    cc_group(keys, vals, label.data(), thr_of.data(), n, *res);
    *n_comp = res->weight.size(); *n_keys = res->keys.size(); *levels = lv; *result = res;
    return 0;
}

extern "C" void cc_emu_fetch(void *result, uint64_t *off, long long *keys, long long *weight, int32_t *thr) {
    CcResult *res = (CcResult *)result;
    memcpy(off, res->off.data(), res->off.size() * 8);
    if (!res->keys.empty()) memcpy(keys, res->keys.data(), res->keys.size() * 8);
    if (!res->weight.empty()) { memcpy(weight, res->weight.data(), res->weight.size() * 8); memcpy(thr, res->thr.data(), res->thr.size() * 4); }
    delete res;
}
