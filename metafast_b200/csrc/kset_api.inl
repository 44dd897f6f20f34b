// kset_api.inl -- C ABI of the device-side (k-mer -> short) maps (include/mfkc.h, "set algebra"); included by mfkc.cu.
// Kernels: kset.cuh.  Every function cites the reference lines it replaces in include/mfkc.h.

struct SeqResult {                           // result of mfkc_kset_sequences_begin, handed out by _fetch
    std::vector<mfkc::SeqRecord> recs;       // accepted sequences in output order
    std::vector<unsigned long long> off;     // base offsets, n + 1 entries
    char *d_bases = nullptr;
};

struct mfkc_kset {
    mfkc_ctx *ctx = nullptr;
    unsigned long long *keys = nullptr; uint32_t *vals = nullptr; uint64_t n = 0;     // finalized: ascending unique keys
    unsigned long long *pk = nullptr; uint32_t *pv = nullptr; uint64_t p_cap = 0, p_ub = 0;   // records loaded, not merged yet
    unsigned long long *d_cursor = nullptr;
    uint8_t *sel = nullptr; uint64_t sel_n = 0, sel_cursor = 0; bool sel_valid = false;
    SeqResult *seq = nullptr;                // no global state: the pending sequences belong to the map
    mfkc::CcResult *comps = nullptr;         // pending components (mfkc_kset_components_begin -> _fetch)
};

static void kset_drop_sequences(mfkc_kset *ks) {
    if (ks->seq) { cudaFree(ks->seq->d_bases); delete ks->seq; ks->seq = nullptr; }
}
static void kset_drop_components(mfkc_kset *ks) { delete ks->comps; ks->comps = nullptr; }

static int kset_block_scan(mfkc_ctx *ctx, cudaStream_t st, unsigned long long *d_blk, int grid, unsigned long long *total) {
    std::vector<unsigned long long> h(grid);
    CU_TRY(cudaMemcpyAsync(h.data(), d_blk, (size_t)grid * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    unsigned long long run = 0;
    for (int i = 0; i < grid; i++) { const unsigned long long c = h[i]; h[i] = run; run += c; }
    CU_TRY(cudaMemcpyAsync(d_blk, h.data(), (size_t)grid * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));
    *total = run;
    return MFKC_OK;
}

extern "C" int mfkc_kset_create(mfkc_ctx *ctx, mfkc_kset **out) {
    if (!ctx || !out) return MFKC_E_BADARG;
    if (ctx->k128) return fail(ctx, MFKC_E_BADARG, "the k-mer set tools work on 64-bit keys (k <= 31), like the reference");
    CU_TRY(cudaSetDevice(ctx->device));
    mfkc_kset *ks = new mfkc_kset();
    ks->ctx = ctx;
    if (cudaMalloc(&ks->d_cursor, sizeof(unsigned long long)) != cudaSuccess) { delete ks; return fail(ctx, MFKC_E_OOM, "cudaMalloc"); }
    cudaMemsetAsync(ks->d_cursor, 0, sizeof(unsigned long long), ctx->compute);
    *out = ks;
    return MFKC_OK;
}

extern "C" void mfkc_kset_destroy(mfkc_kset *ks) {
    if (!ks) return;
    cudaSetDevice(ks->ctx->device);
    cudaStreamSynchronize(ks->ctx->compute);
    cudaFree(ks->keys); cudaFree(ks->vals); cudaFree(ks->pk); cudaFree(ks->pv); cudaFree(ks->d_cursor); cudaFree(ks->sel);
    kset_drop_sequences(ks);
    kset_drop_components(ks);
    delete ks;
}

extern "C" int mfkc_kset_load_records(mfkc_kset *ks, const uint8_t *be_records, uint64_t n_records, int32_t freq_threshold) {
    if (!ks || (!be_records && n_records)) return MFKC_E_BADARG;
    mfkc_ctx *ctx = ks->ctx;
    if (ks->n) return fail(ctx, MFKC_E_STATE, "mfkc_kset_load_records on a map that is already finished");
    if (n_records == 0) return MFKC_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    if (ks->p_ub + n_records > ks->p_cap) {
        const uint64_t cap = std::max<uint64_t>((ks->p_ub + n_records) * 3 / 2, 1u << 20);
        unsigned long long *nk = nullptr; uint32_t *nv = nullptr;
        CU_TRY(cudaStreamSynchronize(st));
        if (big_alloc(ctx, (void **)&nk, cap * 8) != cudaSuccess || big_alloc(ctx, (void **)&nv, cap * 4) != cudaSuccess) { cudaFree(nk); return fail(ctx, MFKC_E_OOM, "cannot allocate the k-mer record buffer"); }
        if (ks->p_ub) {
            CU_TRY(cudaMemcpy(nk, ks->pk, ks->p_ub * 8, cudaMemcpyDeviceToDevice));
            CU_TRY(cudaMemcpy(nv, ks->pv, ks->p_ub * 4, cudaMemcpyDeviceToDevice));
        }
        cudaFree(ks->pk); cudaFree(ks->pv);
        ks->pk = nk; ks->pv = nv; ks->p_cap = cap;
    }
    Staging &s = ctx->st[0];
    CU_TRY(cudaStreamSynchronize(st));                 // the previous chunk's parse kernel has read the staging buffer
    TRY(stage_records(ctx, s, be_records, n_records));
    CU_TRY(cudaStreamSynchronize(s.stream));
    kset_parse_kernel<<<grid_for(ctx, n_records, 256, 8), 256, 0, st>>>(s.d_bases, n_records, freq_threshold, ks->pk, ks->pv, ks->d_cursor);
    CU_TRY(cudaGetLastError());
    ks->p_ub += n_records;
    ks->sel_valid = false;
    return MFKC_OK;
}

extern "C" int mfkc_kset_load_finish(mfkc_kset *ks) {
    if (!ks) return MFKC_E_BADARG;
    mfkc_ctx *ctx = ks->ctx;
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    unsigned long long pn = 0;
    CU_TRY(cudaMemcpyAsync(&pn, ks->d_cursor, sizeof pn, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (pn) {
        unsigned long long *alt = nullptr; uint32_t *valt = nullptr;
        TMP_ALLOC(alt, pn * 8); TMP_ALLOC(valt, pn * 4);
        unsigned long long *a = ks->pk, *b = alt; uint32_t *va = ks->pv, *vb = valt;
        int r = radix_sort<uint32_t, true>(ctx, st, a, b, va, vb, pn, 64);
        unsigned long long *ok = nullptr; uint32_t *oc = nullptr; uint64_t on = 0;
        if (r == MFKC_OK) r = rle_sorted(ctx, st, a, va, pn, &ok, &oc, &on);       // addAndBound of positive values = clamped sum
        TMP_FREE(alt); TMP_FREE(valt);
        if (r != MFKC_OK) return r;
        // move the result out of the pool into plain allocations owned by the map
        cudaFree(ks->keys); cudaFree(ks->vals); ks->keys = nullptr; ks->vals = nullptr;
        if (big_alloc(ctx, (void **)&ks->keys, std::max<uint64_t>(on, 1) * 8) != cudaSuccess || big_alloc(ctx, (void **)&ks->vals, std::max<uint64_t>(on, 1) * 4) != cudaSuccess)
            return fail(ctx, MFKC_E_OOM, "cannot allocate the k-mer map");
        CU_TRY(cudaMemcpyAsync(ks->keys, ok, on * 8, cudaMemcpyDeviceToDevice, st));
        CU_TRY(cudaMemcpyAsync(ks->vals, oc, on * 4, cudaMemcpyDeviceToDevice, st));
        CU_TRY(cudaStreamSynchronize(st));
        TMP_FREE(ok); TMP_FREE(oc);
        ks->n = on;
    }
    cudaFree(ks->pk); cudaFree(ks->pv); ks->pk = nullptr; ks->pv = nullptr; ks->p_cap = ks->p_ub = 0;
    CU_TRY(cudaMemsetAsync(ks->d_cursor, 0, sizeof(unsigned long long), ctx->compute));
    ks->sel_valid = false;
    return MFKC_OK;
}

extern "C" int mfkc_kset_size(mfkc_kset *ks, uint64_t *n) {
    if (!ks || !n) return MFKC_E_BADARG;
    *n = ks->n;
    return MFKC_OK;
}

extern "C" int mfkc_kset_reset_values(mfkc_kset *ks) {
    if (!ks) return MFKC_E_BADARG;
    mfkc_ctx *ctx = ks->ctx;
    CU_TRY(cudaSetDevice(ctx->device));
    if (ks->n) kset_fill_kernel<<<grid_for(ctx, ks->n, 256, 8), 256, 0, ctx->compute>>>(ks->vals, ks->n, 0u);
    CU_TRY(cudaGetLastError());
    ks->sel_valid = false;
    return MFKC_OK;
}

// flags + order-preserving compaction helper: entries of `src` with value > thr (and the filter rule) -> out arrays
template <typename V>
static int kset_compact(mfkc_ctx *ctx, const mfkc_kset *src, int thr, const mfkc_kset *filter, int fthr, uint32_t tag,
                        unsigned long long *out_keys, V *out_vals, uint64_t out_cap, uint64_t *n_out) {
    *n_out = 0;
    if (!src->n) return MFKC_OK;
    cudaStream_t st = ctx->compute;
    uint8_t *flags = nullptr; unsigned long long *d_blk = nullptr;
    const int grid = grid_for(ctx, src->n, 256, 8);
    TMP_ALLOC(flags, src->n); TMP_ALLOC(d_blk, (size_t)grid * 8);
    kset_flag_kernel<<<grid, 256, 0, st>>>(src->keys, src->vals, src->n, thr, filter ? filter->keys : nullptr, filter ? filter->vals : nullptr,
                                           filter ? filter->n : 0, filter ? 1 : 0, fthr, flags);
    kset_flag_count_kernel<<<grid, 256, 0, st>>>(flags, src->n, d_blk);
    CU_TRY(cudaGetLastError());
    unsigned long long total = 0;
    TRY(kset_block_scan(ctx, st, d_blk, grid, &total));
    if (out_keys && total <= out_cap && total)
        kset_flag_write_kernel<V><<<grid, 256, 0, st>>>(src->keys, src->vals, flags, src->n, d_blk, out_keys, out_vals, tag);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(st));
    TMP_FREE(flags); TMP_FREE(d_blk);
    *n_out = total;
    if (out_keys && total > out_cap) return fail(ctx, MFKC_E_STATE, "internal: compaction buffer too small");
    return MFKC_OK;
}

extern "C" int mfkc_kset_update(mfkc_kset *dst, const mfkc_kset *src, int op, int32_t thr) {
    if (!dst || !src || dst->ctx != src->ctx || dst == src) return MFKC_E_BADARG;
    mfkc_ctx *ctx = dst->ctx;
    if (op != KSET_ADD && op != KSET_INC && op != KSET_ZERO) return fail(ctx, MFKC_E_BADARG, "unknown map update");
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    dst->sel_valid = false;
    if (!src->n) return MFKC_OK;
    if (op == KSET_ZERO) {
        if (dst->n) kset_zero_kernel<<<grid_for(ctx, dst->n, 256, 8), 256, 0, st>>>(dst->keys, dst->vals, dst->n, src->keys, src->vals, src->n, thr);
        CU_TRY(cudaGetLastError());
        return MFKC_OK;
    }
    // union: (dst entries, then the src entries with value > thr, tagged) -> stable sort -> pair combine
    const uint64_t m_cap = dst->n + src->n;
    unsigned long long *k1 = nullptr, *k2 = nullptr; uint32_t *v1 = nullptr, *v2 = nullptr;
    TMP_ALLOC(k1, m_cap * 8); TMP_ALLOC(k2, m_cap * 8); TMP_ALLOC(v1, m_cap * 4); TMP_ALLOC(v2, m_cap * 4);
    if (dst->n) {
        CU_TRY(cudaMemcpyAsync(k1, dst->keys, dst->n * 8, cudaMemcpyDeviceToDevice, st));
        CU_TRY(cudaMemcpyAsync(v1, dst->vals, dst->n * 4, cudaMemcpyDeviceToDevice, st));
        kset_tag_kernel<<<grid_for(ctx, dst->n, 256, 8), 256, 0, st>>>(v1, dst->n, 0u);
    }
    uint64_t ns = 0;
    int r = kset_compact<uint32_t>(ctx, src, thr, nullptr, 0, 1u << 16, k1 + dst->n, v1 + dst->n, src->n, &ns);
    const uint64_t m = dst->n + ns;
    unsigned long long *ok = nullptr; uint32_t *ov = nullptr; unsigned long long total = 0;
    if (r == MFKC_OK && ns) {
        r = radix_sort<uint32_t, true>(ctx, st, k1, k2, v1, v2, m, 64);
        if (r == MFKC_OK) {
            const int grid = grid_for(ctx, m, 256, 8);
            unsigned long long *d_blk = nullptr;
            TMP_ALLOC(d_blk, (size_t)grid * 8);
            rle_mark_kernel<<<grid, 256, 0, st>>>(k1, m, d_blk);
            r = kset_block_scan(ctx, st, d_blk, grid, &total);
            if (r == MFKC_OK) {
                if (big_alloc(ctx, (void **)&ok, std::max<uint64_t>(total, 1) * 8) != cudaSuccess || big_alloc(ctx, (void **)&ov, std::max<uint64_t>(total, 1) * 4) != cudaSuccess) {
                    cudaFree(ok); ok = nullptr; r = fail(ctx, MFKC_E_OOM, "cannot allocate the k-mer map");
                } else {
                    kset_combine_kernel<<<grid, 256, 0, st>>>(k1, v1, m, op, d_blk, ok, ov);
                    if (cudaStreamSynchronize(st) != cudaSuccess) r = fail(ctx, MFKC_E_CUDA, "map update kernel failed");
                }
            }
            TMP_FREE(d_blk);
        }
    }
    TMP_FREE(k1); TMP_FREE(k2); TMP_FREE(v1); TMP_FREE(v2);
    if (r != MFKC_OK) { cudaFree(ok); cudaFree(ov); return r; }
    if (ns) {
        cudaFree(dst->keys); cudaFree(dst->vals);
        dst->keys = ok; dst->vals = ov; dst->n = total;
    }
    return MFKC_OK;
}

extern "C" int mfkc_kset_select_begin(mfkc_kset *hm, const mfkc_kset *filter, int32_t threshold, int32_t filter_threshold, uint64_t *n_good) {
    if (!hm || (filter && filter->ctx != hm->ctx)) return MFKC_E_BADARG;
    mfkc_ctx *ctx = hm->ctx;
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    cudaFree(hm->sel); hm->sel = nullptr; hm->sel_n = hm->sel_cursor = 0; hm->sel_valid = false;
    uint64_t good = 0;
    if (hm->n) {
        unsigned long long *k1 = nullptr; uint16_t *c1 = nullptr;
        TMP_ALLOC(k1, hm->n * 8); TMP_ALLOC(c1, hm->n * 2);
        int r = kset_compact<uint16_t>(ctx, hm, threshold, filter, filter_threshold, 0u, k1, c1, hm->n, &good);
        if (r == MFKC_OK && good) {
            if (big_alloc(ctx, (void **)&hm->sel, good * 10) != cudaSuccess) r = fail(ctx, MFKC_E_OOM, "cannot allocate the record buffer");
            else {
                records_kernel<<<grid_for(ctx, good, 256, 8), 256, 0, st>>>(k1, c1, good, reinterpret_cast<uint16_t *>(hm->sel));
                if (cudaStreamSynchronize(st) != cudaSuccess) r = fail(ctx, MFKC_E_CUDA, "records kernel failed");
            }
        }
        TMP_FREE(k1); TMP_FREE(c1);
        if (r != MFKC_OK) return r;
    }
    hm->sel_n = good; hm->sel_valid = true;
    if (n_good) *n_good = good;
    return MFKC_OK;
}

extern "C" int mfkc_kset_select_next(mfkc_kset *hm, uint8_t *out, size_t cap, size_t *written) {
    if (!hm || !written) return MFKC_E_BADARG;
    mfkc_ctx *ctx = hm->ctx;
    if (!hm->sel_valid) return fail(ctx, MFKC_E_STATE, "mfkc_kset_select_next without mfkc_kset_select_begin");
    CU_TRY(cudaSetDevice(ctx->device));
    const uint64_t take = std::min<uint64_t>(hm->sel_n - hm->sel_cursor, cap / 10);
    if (take && !out) return MFKC_E_BADARG;
    if (take) CU_TRY(cudaMemcpy(out, hm->sel + hm->sel_cursor * 10, (size_t)take * 10, cudaMemcpyDeviceToHost));
    hm->sel_cursor += take;
    *written = (size_t)take * 10;
    return MFKC_OK;
}

extern "C" int mfkc_kset_histogram(mfkc_kset *ks, uint64_t hist[MFKC_HIST_BINS]) {
    if (!ks || !hist) return MFKC_E_BADARG;
    mfkc_ctx *ctx = ks->ctx;
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    ctx->hist_valid = false;                              // the context's histogram buffer is borrowed
    CU_TRY(cudaMemsetAsync(ctx->d_hist, 0, MFKC_HIST_BINS * sizeof(unsigned long long), st));
    if (ks->n) kset_hist_kernel<<<grid_for(ctx, ks->n, 256, 8), 256, 0, st>>>(ks->vals, ks->n, ctx->d_hist);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(hist, ctx->d_hist, MFKC_HIST_BINS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return MFKC_OK;
}

// ---- seq-builder on a k-mer map (kernels at the end of kset.cuh) ----------------------------------
extern "C" int mfkc_kset_sequences_begin(mfkc_kset *hm, int32_t freq_threshold, int32_t len_threshold, uint64_t *n_sequences, uint64_t *n_bases) {
    if (!hm || !n_sequences || !n_bases) return MFKC_E_BADARG;
    mfkc_ctx *ctx = hm->ctx;
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    SeqResult res;
    kset_drop_sequences(hm);
    *n_sequences = 0; *n_bases = 0;
    res.off.assign(1, 0);
    if (hm->n) {
        Slot *tab = nullptr;
        const uint64_t cap = hm->n * 2 + 64;
        TRY(table_alloc(ctx, cap, &tab, true));
        seq_index_build_kernel<<<grid_for(ctx, hm->n, 256, 8), 256, 0, st>>>(hm->keys, hm->vals, hm->n, tab, cap);
        SeqIndex ix; ix.tab = tab; ix.cap = cap; ix.k = ctx->cfg.k; ix.thr = freq_threshold;
        unsigned long long *d_cur = nullptr; unsigned long long h_cur[2] = {0, 0};
        TMP_ALLOC(d_cur, 2 * sizeof(unsigned long long));
        CU_TRY(cudaMemsetAsync(d_cur, 0, 2 * sizeof(unsigned long long), st));
        const uint32_t len_thr = len_threshold < 0 ? 0u : (uint32_t)len_threshold;
        seq_find_kernel<<<grid_for(ctx, hm->n, 256, 8), 256, 0, st>>>(hm->keys, hm->vals, hm->n, ix, len_thr, nullptr, d_cur);      // count
        CU_TRY(cudaMemcpyAsync(h_cur, d_cur, sizeof h_cur, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        const uint64_t ns = h_cur[0], nb = h_cur[1];
        int r = MFKC_OK;
        if (ns) {
            SeqRecord *d_recs = nullptr; unsigned long long *d_off = nullptr;
            TMP_ALLOC(d_recs, ns * sizeof(SeqRecord));
            TMP_ALLOC(d_off, ns * sizeof(unsigned long long));
            CU_TRY(cudaMemsetAsync(d_cur, 0, 2 * sizeof(unsigned long long), st));
            seq_find_kernel<<<grid_for(ctx, hm->n, 256, 8), 256, 0, st>>>(hm->keys, hm->vals, hm->n, ix, len_thr, d_recs, d_cur);   // records
            res.recs.resize(ns);
            CU_TRY(cudaMemcpyAsync(res.recs.data(), d_recs, ns * sizeof(SeqRecord), cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
            // deterministic order: ascending (start key, orientation); the reference's order is thread timing
            std::sort(res.recs.begin(), res.recs.end(), [](const SeqRecord &a, const SeqRecord &b) { return a.order < b.order; });
            res.off.resize(ns + 1);
            for (uint64_t i = 0; i < ns; i++) res.off[i + 1] = res.off[i] + res.recs[i].length;
            if (res.off[ns] != nb) r = fail(ctx, MFKC_E_STATE, "internal: sequence length mismatch");
            if (r == MFKC_OK && big_alloc(ctx, (void **)&res.d_bases, nb) != cudaSuccess) r = fail(ctx, MFKC_E_OOM, "cannot allocate the sequence buffer");
            if (r == MFKC_OK) {
                CU_TRY(cudaMemcpyAsync(d_recs, res.recs.data(), ns * sizeof(SeqRecord), cudaMemcpyHostToDevice, st));
                CU_TRY(cudaMemcpyAsync(d_off, res.off.data(), ns * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
                seq_write_kernel<<<grid_for(ctx, ns, 128, 16), 128, 0, st>>>(d_recs, d_off, ns, ix, res.d_bases);
                if (cudaStreamSynchronize(st) != cudaSuccess) r = fail(ctx, MFKC_E_CUDA, "sequence kernel failed");
            }
            TMP_FREE(d_recs); TMP_FREE(d_off);
        }
        TMP_FREE(d_cur);
        CU_TRY(cudaStreamSynchronize(st));
        cudaFree(tab);
        if (r != MFKC_OK) { cudaFree(res.d_bases); return r; }
        *n_sequences = ns; *n_bases = nb;
    }
    hm->seq = new SeqResult(std::move(res));
    return MFKC_OK;
}

extern "C" int mfkc_kset_sequences_fetch(mfkc_kset *hm, uint64_t *offsets, char *bases, uint32_t *av_weight, uint32_t *min_weight, uint32_t *max_weight) {
    if (!hm || !offsets) return MFKC_E_BADARG;
    mfkc_ctx *ctx = hm->ctx;
    if (!hm->seq) return fail(ctx, MFKC_E_STATE, "mfkc_kset_sequences_fetch without mfkc_kset_sequences_begin");
    SeqResult res = std::move(*hm->seq);
    delete hm->seq; hm->seq = nullptr;
    CU_TRY(cudaSetDevice(ctx->device));
    const uint64_t ns = res.recs.size();
    for (uint64_t i = 0; i <= ns; i++) offsets[i] = res.off[i];
    for (uint64_t i = 0; i < ns; i++) {
        if (av_weight) av_weight[i] = res.recs[i].av_weight;
        if (min_weight) min_weight[i] = res.recs[i].min_weight;
        if (max_weight) max_weight[i] = res.recs[i].max_weight;
    }
    cudaError_t e = cudaSuccess;
    if (ns && bases) e = cudaMemcpy(bases, res.d_bases, res.off[ns], cudaMemcpyDeviceToHost);
    cudaFree(res.d_bases);
    if (e != cudaSuccess) return fail(ctx, MFKC_E_CUDA, "copy of the sequences failed");
    return MFKC_OK;
}

// ---- component-cutter's graph half on a k-mer map (kernels: components.cuh, grouping: components_host.h) --------------
// Device buffers of one run; freed on every exit path
struct CcBuffers {
    Slot *tab = nullptr; uint8_t *active = nullptr; uint32_t *label = nullptr, *thr_of = nullptr, *parent = nullptr, *size = nullptr;
    unsigned long long *counters = nullptr;
    ~CcBuffers() { cudaFree(tab); cudaFree(active); cudaFree(label); cudaFree(thr_of); cudaFree(parent); cudaFree(size); cudaFree(counters); }
};

extern "C" int mfkc_kset_components_begin(mfkc_kset *hm, int64_t min_component_size, int64_t max_component_size, uint64_t *n_components,
                                          uint64_t *n_kmers) {
    if (!hm || !n_components || !n_kmers) return MFKC_E_BADARG;
    mfkc_ctx *ctx = hm->ctx;
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    kset_drop_components(hm);
    *n_components = 0; *n_kmers = 0;
    const uint64_t n = hm->n;
    if (n >= 0xFFFFFFFFull) return fail(ctx, MFKC_E_BADARG, "mfkc_kset_components: the map has more than 2^32 - 2 entries");
    mfkc::CcResult *res = new mfkc::CcResult();
    res->off.assign(1, 0);
    if (n) {
        const int k = ctx->cfg.k;
        const uint64_t cap = n * 2 + 64;
        CcBuffers b;
        bool oom = big_alloc(ctx, (void **)&b.tab, cap * sizeof(Slot)) != cudaSuccess;
        oom = oom || big_alloc(ctx, (void **)&b.active, n) != cudaSuccess;
        oom = oom || big_alloc(ctx, (void **)&b.label, n * 4) != cudaSuccess || big_alloc(ctx, (void **)&b.thr_of, n * 4) != cudaSuccess;
        oom = oom || big_alloc(ctx, (void **)&b.parent, n * 4) != cudaSuccess || big_alloc(ctx, (void **)&b.size, n * 4) != cudaSuccess;
        oom = oom || cudaMalloc((void **)&b.counters, 2 * sizeof(unsigned long long)) != cudaSuccess;
        if (oom) { delete res; return fail(ctx, MFKC_E_OOM, "cannot allocate the component buffers"); }
        const int grid = grid_for(ctx, n, 256, 8);
        cudaError_t e = cudaMemsetAsync(b.tab, 0xFF, cap * sizeof(Slot), st);                   // every key = EMPTY_KEY
        cc_index_build_kernel<<<grid, 256, 0, st>>>(hm->keys, n, b.tab, cap);
        cc_begin_kernel<<<grid, 256, 0, st>>>(hm->vals, n, b.active, b.label, b.thr_of);
        CcIndex ix; ix.tab = b.tab; ix.cap = cap;
        // one level per frequency threshold (Task.run: curFreqThreshold = usedFreqThreshold + 1); values are shorts, so
        // nothing can stay active past 32767
        for (int thr = 1; thr <= 32767 && e == cudaSuccess; thr++) {
            unsigned long long counters[2] = {0, 0};
            e = cudaMemsetAsync(b.counters, 0, sizeof counters, st);
            cc_level_init_kernel<<<grid, 256, 0, st>>>(n, b.parent, b.size);
            cc_union_kernel<<<grid, 256, 0, st>>>(hm->keys, n, b.active, ix, k, b.parent);
            cc_count_kernel<<<grid, 256, 0, st>>>(n, b.active, b.parent, b.size);
            cc_classify_kernel<<<grid, 256, 0, st>>>(hm->vals, n, b.active, b.parent, b.size, (long long)min_component_size,
                                                      (long long)max_component_size, thr, b.label, b.thr_of, b.counters);
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(counters, b.counters, sizeof counters, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess || !counters[0]) break;
        }
        std::vector<unsigned long long> h_keys(n);
        std::vector<uint32_t> h_vals(n), h_label(n), h_thr(n);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_keys.data(), hm->keys, n * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_vals.data(), hm->vals, n * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_label.data(), b.label, n * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_thr.data(), b.thr_of, n * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { delete res; return fail(ctx, MFKC_E_CUDA, cudaGetErrorString(e)); }
        mfkc::cc_group(h_keys.data(), h_vals.data(), h_label.data(), h_thr.data(), n, *res);
    }
    *n_components = res->weight.size(); *n_kmers = res->keys.size();
    hm->comps = res;
    return MFKC_OK;
}

extern "C" int mfkc_kset_components_fetch(mfkc_kset *hm, uint64_t *comp_offsets, int64_t *keys, int64_t *weights, int32_t *thresholds) {
    if (!hm || !comp_offsets) return MFKC_E_BADARG;
    mfkc_ctx *ctx = hm->ctx;
    if (!hm->comps) return fail(ctx, MFKC_E_STATE, "mfkc_kset_components_fetch without mfkc_kset_components_begin");
    const mfkc::CcResult &res = *hm->comps;
    const size_t nc = res.weight.size();
    if (!keys && !res.keys.empty()) return MFKC_E_BADARG;
    memcpy(comp_offsets, res.off.data(), (nc + 1) * sizeof(uint64_t));
    if (!res.keys.empty()) memcpy(keys, res.keys.data(), res.keys.size() * sizeof(int64_t));
    if (weights && nc) memcpy(weights, res.weight.data(), nc * sizeof(int64_t));
    if (thresholds && nc) memcpy(thresholds, res.thr.data(), nc * sizeof(int32_t));
    kset_drop_components(hm);
    return MFKC_OK;
}
