// LINE trainer (included by train_line_f32.cu / train_line_f64.cu)
#include "host_common.h"

template <typename T>
int train_line_t(smore_model_s* m, const smore_train_params* p) {
    const int vtab = 0, ctab = p->order == 1 ? 0 : 1;
    return dispatch_dim<T>(m->dim, [&](auto cfg) -> int {
        using C = decltype(cfg);
        const bool cpp = p->semantics == SMORE_SEM_CPP;
        const int shard = m->g->world == 1 ? 0 : (m->replica[vtab] ? 2 : 1);
        void (*kern)(TrainArgs<T>) =
            cpp ? (shard == 2 ? k_line<C, false, 2> : shard == 1 ? k_line<C, false, 1> : k_line<C, false, 0>)
                : (shard == 2 ? k_line<C, true, 2> : shard == 1 ? k_line<C, true, 1> : k_line<C, true, 0>);
        const size_t smem = batch_smem_bytes<T>(m->g->world > 1 ? 2 : cpp ? 0 : 1, p->negative_samples,
                                                shard == 1 ? C::EPL * 32 : 0);
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        Launch L;
        if (p->mode == SMORE_MODE_DETERMINISTIC) { L.blocks = 1; L.warps = 1; }
        else if (int rc = pick_grid(kern, smem, p->max_warps, p->total, L)) return rc;
        // row-sharded: this rank draws sources from its own vertices only and runs its share of the global total
        const uint64_t total_local = m->g->world == 1 ? p->total : (uint64_t)llround((double)p->total * m->g->src_mass_frac);
        // jobs = total / workers (LINE.cpp:124); the C++ loop starts count at 1 and runs while count < jobs
        const uint64_t jobs = total_local / (uint64_t)L.warps;
        const uint64_t trips = cpp ? (jobs > 0 ? jobs - 1 : 0) : jobs;
        if (int rc = init_state(m, L.warps, cpp ? 1 : 0, p->alpha, p)) return rc;
        TrainArgs<T> a = base_args<T>(m, p, L.warps, (double)total_local, cpp ? 1 : 0, vtab, ctab,
                                      m->g->world == 1 ? 1.0 : m->g->src_mass_frac);
        a.jobs = trips;
        a.replica_v = (T*)m->replica[vtab];
        Timer t;
        if (int rc = t.start()) return rc;
        kern<<<L.blocks, kBlockThreads, smem>>>(a);
        g_launches++;
        CU(cudaGetLastError());
        if (int rc = t.stop(&m->st_ms)) return rc;
        m->st_samples = trips * (uint64_t)L.warps;
        return collect_stats(m, L.warps);
    });
}


// ---- bulk-exchange mode (exchange_host.h; device side: batch_kernels.cuh) --------------------------------------------
// Every rank runs the same number of super-batches (the collectives must match): nsb = ceil(total / (superbatch*world)),
// and splits its own share of the samples evenly over them.
template <typename T>
int train_line_exchange_t(smore_model_s** ms, int n, const smore_train_params* p, ExchTransport& tr) {
    smore_model_s* m0 = ms[0];
    const int vtab = 0, ctab = p->order == 1 ? 0 : 1;
    return dispatch_dim<T>(m0->dim, [&](auto cfg) -> int {
        using C = decltype(cfg);
        const bool cpp = p->semantics == SMORE_SEM_CPP;
        const int world = m0->g->world;
        void (*kern)(TrainArgs<T>) = cpp ? k_line<C, false, 3> : k_line<C, true, 3>;
        const size_t smem = batch_smem_bytes<T>(2, p->negative_samples, 0);
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        Launch L;
        if (p->mode == SMORE_MODE_DETERMINISTIC) { L.blocks = 1; L.warps = 1; }
        else if (int rc = pick_grid(kern, smem, p->max_warps, p->total, L)) return rc;
        const size_t row_bytes = (size_t)m0->dim * sizeof(T);
        const uint64_t per_sb = (uint64_t)m0->xch->superbatch * (uint64_t)world;
        const uint64_t nsb = std::max<uint64_t>(1, (p->total + per_sb - 1) / per_sb);
        struct Shard {
            TrainArgs<T> a;
            uint64_t trips, jobs_sb, done;
        };
        std::vector<Shard> sh((size_t)n);
        for (int i = 0; i < n; ++i) {
            smore_model_s* m = ms[i];
            smore_exchange_s* x = m->xch;
            const uint64_t total_local = (uint64_t)llround((double)p->total * m->g->src_mass_frac);
            const uint64_t jobs = total_local / (uint64_t)L.warps;
            Shard& s = sh[(size_t)i];
            s.trips = cpp ? (jobs > 0 ? jobs - 1 : 0) : jobs;
            s.jobs_sb = (s.trips + nsb - 1) / nsb;
            s.done = 0;
            if (int rc = init_state(m, L.warps, cpp ? 1 : 0, p->alpha, p)) return rc;
            s.a = base_args<T>(m, p, L.warps, (double)total_local, cpp ? 1 : 0, vtab, ctab, m->g->src_mass_frac);
            if (n > 1) s.a.stream_base += (uint64_t)m->g->rank << 20;  // one call drives all the shards: disjoint streams
            // per-super-batch buffers on the requester side
            const uint64_t sb_samples = std::max<uint64_t>(1, s.jobs_sb * (uint64_t)L.warps);
            uint64_t hsize = 1024;
            while (hsize < 2 * sb_samples) hsize <<= 1;
            if (hsize > (1ull << 30)) return fail(SMORE_E_INVALID, "super-batch too large");
            x->req_stride = (int64_t)sb_samples;
            if (int rc = x->hkey.ensure(hsize * 4)) return rc;
            if (int rc = x->hval.ensure(hsize * 4)) return rc;
            if (int rc = x->req.ensure((size_t)world * sb_samples * 4)) return rc;
            if (int rc = x->cnt.ensure(2 * kMaxWorld * 4)) return rc;
            if (int rc = x->off.ensure(kMaxWorld * 4)) return rc;
            s.a.x = ExchDev{(int32_t*)x->hkey.p, (int32_t*)x->hval.p, (uint32_t)(hsize - 1), (int32_t*)x->req.p,
                            x->req_stride, (int32_t*)x->cnt.p, (const int32_t*)x->off.p, nullptr,
                            x->n_hot ? (const uint32_t*)x->hot.p : nullptr};
            x->st_rows_moved = 0;
            x->st_superbatches = nsb;
        }
        int dev = 0, sms = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int row_blocks = sms * 8;
        Timer t;
        if (int rc = t.start()) return rc;
        for (uint64_t sb = 0; sb < nsb; ++sb) {
            // 1. which remote vertex rows will this super-batch touch?
            for (int i = 0; i < n; ++i) {
                Shard& s = sh[(size_t)i];
                smore_exchange_s* x = ms[i]->xch;
                s.a.jobs = std::min<uint64_t>(s.jobs_sb, s.trips - s.done);
                CU(cudaMemsetAsync(x->hkey.p, 0xff, ((size_t)s.a.x.hmask + 1) * 4, 0));
                CU(cudaMemsetAsync(x->cnt.p, 0, 2 * kMaxWorld * 4, 0));
                if (s.a.jobs) {
                    k_line_requests<<<L.blocks, kBlockThreads>>>(s.a.g, s.a.x, s.a.state, L.warps, s.a.jobs, s.a.seed,
                                                                s.a.stream_base, s.a.K);
                    g_launches++;
                }
            }
            CU(cudaGetLastError());
            if (int rc = tr.counts(ms, n)) return rc;
            for (int i = 0; i < n; ++i) {
                smore_exchange_s* x = ms[i]->xch;
                const size_t n_out = (size_t)x->off_out[kMaxWorld], n_in = (size_t)x->off_in[kMaxWorld];
                if (int rc = x->wrk.ensure(std::max<size_t>(n_out, 1) * row_bytes)) return rc;
                if (int rc = x->req_in.ensure(std::max<size_t>(n_in, 1) * 4)) return rc;
                if (int rc = x->sent.ensure(std::max<size_t>(n_in, 1) * row_bytes)) return rc;
                if (int rc = x->back.ensure(std::max<size_t>(n_in, 1) * row_bytes)) return rc;
                int32_t off32[kMaxWorld];
                for (int r = 0; r < kMaxWorld; ++r) off32[r] = (int32_t)x->off_out[r];
                CU(cudaMemcpyAsync(x->off.p, off32, sizeof(off32), cudaMemcpyHostToDevice, 0));
                sh[(size_t)i].a.x.wrk = x->wrk.p;
                x->st_rows_moved += n_out;
            }
            // 2. request lists to the owners; owners gather the rows; rows to the requesters
            if (int rc = tr.a2a(ms, n, ExchTransport::REQ, row_bytes)) return rc;
            for (int i = 0; i < n; ++i) {
                smore_exchange_s* x = ms[i]->xch;
                const int64_t n_in = x->off_in[kMaxWorld];
                if (n_in) {
                    k_gather_rows<C><<<row_blocks, kBlockThreads>>>((const T*)ms[i]->tab[vtab], (const int32_t*)x->req_in.p,
                                                                    n_in, (T*)x->sent.p, ms[i]->dim);
                    g_launches++;
                }
            }
            if (int rc = tr.a2a(ms, n, ExchTransport::ROWS_OUT, row_bytes)) return rc;
            // 3. the updates: every row access is local
            for (int i = 0; i < n; ++i) {
                Shard& s = sh[(size_t)i];
                if (s.a.jobs) {
                    kern<<<L.blocks, kBlockThreads, smem>>>(s.a);
                    g_launches++;
                    s.done += s.a.jobs;
                }
            }
            // 4. rows back to their owners, who add what changed
            if (int rc = tr.a2a(ms, n, ExchTransport::ROWS_BACK, row_bytes)) return rc;
            for (int i = 0; i < n; ++i) {
                smore_exchange_s* x = ms[i]->xch;
                const int64_t n_in = x->off_in[kMaxWorld];
                if (n_in) {
                    k_apply_delta<C><<<row_blocks, kBlockThreads>>>((T*)ms[i]->tab[vtab], (const int32_t*)x->req_in.p, n_in,
                                                                    (const T*)x->back.p, (const T*)x->sent.p, ms[i]->dim);
                    g_launches++;
                }
            }
            CU(cudaGetLastError());
        }
        double ms_total = 0;
        if (int rc = t.stop(&ms_total)) return rc;
        for (int i = 0; i < n; ++i) {
            ms[i]->st_ms = ms_total;
            ms[i]->st_samples = sh[(size_t)i].done * (uint64_t)L.warps;
            if (int rc = collect_stats(ms[i], L.warps)) return rc;
        }
        return SMORE_OK;
    });
}
