// components_host.h -- host half of mfkc_kset_components_*: turns the per-k-mer (label, threshold) arrays that the
// device kernels (components.cuh) leave behind into the component list in the reference's order.  Plain C++ (no CUDA),
// shared by mfkc.cu and the host emulation harness tests/emu/cc_emu.cpp.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace mfkc {

struct CcResult {
    std::vector<uint64_t> off;          // n_comp + 1 offsets into keys
    std::vector<long long> keys;        // members, ascending inside a component
    std::vector<long long> weight;      // ConnectedComponent.weight: sum of the members' values
    std::vector<int32_t> thr;           // ConnectedComponent.usedFreqThreshold
};

// label[i] = root entry of the output component of k-mer i or 0xFFFFFFFF, thr_of[i] = the level that placed it.
// Order: ConnectedComponent.compareTo (src/structures/ConnectedComponent.java:125-136: threshold ascending, weight
// descending, size descending) after Collections.sort (src/algo/ComponentsBuilder.java:144); ties (thread / hash-map
// order in the reference) by ascending smallest k-mer.
inline void cc_group(const unsigned long long *keys, const uint32_t *vals, const uint32_t *label, const uint32_t *thr_of, uint64_t n,
                     CcResult &out) {
    struct Comp { unsigned long long first_key; long long weight; uint64_t size; int32_t thr; };
    std::vector<uint32_t> id(n, 0xFFFFFFFFu);               // root entry -> component number
    std::vector<Comp> comps;
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t r = label[i];
        if (r == 0xFFFFFFFFu) continue;
        if (id[r] == 0xFFFFFFFFu) {
            id[r] = (uint32_t)comps.size();
            comps.push_back(Comp{keys[i], 0, 0, (int32_t)thr_of[i]});       // keys ascend: the first member met is the smallest
        }
        Comp &c = comps[id[r]];
        c.size++;
        c.weight += (long long)(short)(vals[i] & 0xFFFFu);                  // comp.add(kmer, value): weight += w
    }
    std::vector<uint32_t> order(comps.size());
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        const Comp &x = comps[a], &y = comps[b];
        if (x.thr != y.thr) return x.thr < y.thr;
        if (x.weight != y.weight) return x.weight > y.weight;
        if (x.size != y.size) return x.size > y.size;
        return x.first_key < y.first_key;
    });
    const size_t nc = comps.size();
    out.off.assign(nc + 1, 0);
    out.weight.resize(nc); out.thr.resize(nc);
    std::vector<uint64_t> cursor(nc);                        // component number -> next free position
    for (size_t j = 0; j < nc; j++) {
        const Comp &c = comps[order[j]];
        out.off[j + 1] = out.off[j] + c.size;
        out.weight[j] = c.weight; out.thr[j] = c.thr;
        cursor[order[j]] = out.off[j];
    }
    out.keys.resize(out.off[nc]);
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t r = label[i];
        if (r != 0xFFFFFFFFu) out.keys[cursor[id[r]]++] = (long long)keys[i];
    }
}

}  // namespace mfkc
