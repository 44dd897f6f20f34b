// cc_emu.cpp -- TEST INFRASTRUCTURE: compiles the component kernels of metafast_b200/csrc/components.cuh for the HOST
// (g++, one emulated CUDA thread: grid 1 x block 1, atomics as plain read-modify-writes) and drives them with the same
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
static const uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
static const dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);
template <class T, class U, class V> static T atomicCAS(T *p, U cmp, V val) { const T old = *p; if (old == (T)cmp) *p = (T)val; return old; }
template <class T, class U> static T atomicAdd(T *p, U v) { const T old = *p; *p = old + (T)v; return old; }

#include "../../metafast_b200/csrc/components.cuh"
#include "../../metafast_b200/csrc/components_host.h"

using namespace mfkc;

extern "C" int cc_emu_components(const unsigned long long *keys, const uint32_t *vals, uint64_t n, int k, long long b1, long long b2,
                                 uint64_t *n_comp, uint64_t *n_keys, int *levels, void **result) {
    std::vector<uint8_t> active(n ? n : 1);
    std::vector<uint32_t> label(n ? n : 1), thr_of(n ? n : 1), parent(n ? n : 1), size(n ? n : 1);
    const uint64_t cap = n * 2 + 64;
    std::vector<Slot> tab(cap);
    memset(tab.data(), 0xFF, cap * sizeof(Slot));
    CcIndex ix; ix.tab = tab.data(); ix.cap = cap;
    int lv = 0;
    if (n) {
        cc_index_build_kernel(keys, n, tab.data(), cap);
        cc_begin_kernel(vals, n, active.data(), label.data(), thr_of.data());
        for (int thr = 1; thr <= 32767; thr++) {
            unsigned long long counters[2] = {0, 0};
            cc_level_init_kernel(n, parent.data(), size.data());
            cc_union_kernel(keys, n, active.data(), ix, k, parent.data());
            cc_count_kernel(n, active.data(), parent.data(), size.data());
            cc_classify_kernel(vals, n, active.data(), parent.data(), size.data(), b1, b2, thr, label.data(), thr_of.data(), counters);
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
