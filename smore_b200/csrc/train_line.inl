// LINE trainer (included by train_line_f32.cu / train_line_f64.cu)
#include "host_common.h"

#include <cuda.h>  // green-context types; the entry points come from cudaGetDriverEntryPoint

template <typename T>
int train_line_t(smore_model_s* m, const smore_train_params* p) {
    const int vtab = 0, ctab = p->order == 1 ? 0 : 1;
    return dispatch_dim<T>(m->dim, [&](auto cfg) -> int {
        using C = decltype(cfg);
        const bool cpp = p->semantics == SMORE_SEM_CPP;
        const int shard = m->g->world == 1 ? 0 : (m->replica[vtab] ? 2 : 1);
        // peer-access mode, LINE-2: split samples on request (neg_mode; the default is the reference's pairing)
        const bool split = shard == 1 && vtab != ctab && p->neg_mode == SMORE_PAIRING_SPLIT;
        // the cp.async row pipeline (batch_kernels.cuh, SMORE_LINE_PIPE) serves the Hogwild C++ LINE-2 loop with K <= 5;
        // everything else runs the same kernel without it (KIND = 4: three resident CTAs instead of two)
        const bool pipe = line_pipe_cfg<C>() && cpp && shard == 0 && vtab != ctab && p->negative_samples + 1 <= kCtxChunk &&
                          p->mode != SMORE_MODE_DETERMINISTIC;
        void (*kern)(TrainArgs<T>) =
            split ? k_line<C, false, 1, 2>
            : cpp ? (shard == 2 ? k_line<C, false, 2> : shard == 1 ? k_line<C, false, 1> : pipe ? k_line<C, false, 0> : k_line<C, false, 0, 4>)
                  : (shard == 2 ? k_line<C, true, 2> : shard == 1 ? k_line<C, true, 1> : k_line<C, true, 0>);
        const size_t smem = batch_smem_bytes<T>(m->g->world > 1 ? (split ? 3 : 2) : cpp ? 0 : 1, p->negative_samples,
                                                shard == 1 ? C::EPL * 32 : 0) + (pipe ? line_pipe_bytes<C>() : 0);
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        Launch L;
        if (p->mode == SMORE_MODE_DETERMINISTIC) { L.blocks = 1; L.warps = 1; }
        else if (int rc = pick_grid(kern, smem, effective_max_warps(p, m->rows * (vtab == ctab ? 1 : 2)), p->total, L)) return rc;
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


// Rotating shards (rotation.cu): one block = the sources of vertex sub-part q (rows resident in `vslot`) x the contexts
// this rank owns. Every row the kernel touches is in local HBM, so this is the single-GPU kernel k_line<.., SHARD = 0> fed
// by the block's edge table: ids are already ROW indices (source: row inside the sub-part, context / negatives: local row
// of the context shard).
template <typename T>
int train_line_block_t(smore_model_s* m, const smore_train_params* p, int q, void* vslot, uint64_t n_samples) {
    return dispatch_dim<T>(m->dim, [&](auto cfg) -> int {
        using C = decltype(cfg);
        const bool cpp = p->semantics == SMORE_SEM_CPP;
        const bool split = p->neg_mode == SMORE_PAIRING_SPLIT;
        const bool pipe = line_pipe_cfg<C>() && cpp && !split && p->negative_samples + 1 <= kCtxChunk && p->mode != SMORE_MODE_DETERMINISTIC;
        void (*kern)(TrainArgs<T>) = split ? k_line<C, false, 0, 2> : cpp ? (pipe ? k_line<C, false, 0> : k_line<C, false, 0, 4>) : k_line<C, true, 0>;
        const size_t smem = batch_smem_bytes<T>(split ? 3 : 2, p->negative_samples, 0) + (pipe ? line_pipe_bytes<C>() : 0);
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        Launch L;
        if (p->mode == SMORE_MODE_DETERMINISTIC) { L.blocks = 1; L.warps = 1; }
        else if (int rc = pick_grid(kern, smem, effective_max_warps(p, m->g->sub_rows[(size_t)q] + m->rows), n_samples, L)) return rc;
        const uint64_t jobs = n_samples / (uint64_t)L.warps;
        if (int rc = init_state(m, L.warps, 0, p->alpha, p)) return rc;
        // p->total = samples of ALL ranks in this episode; the schedule (sched_total / sched_offset) is in the same global unit
        TrainArgs<T> a = base_args<T>(m, p, L.warps, (double)n_samples, 0, 0, 1, p->total ? (double)n_samples / (double)p->total : 1.0);
        const smore_graph_s* g = m->g;
        const int64_t off = g->blk_off[(size_t)q];
        a.g.edge_at = g->d_eat + off;
        a.g.edge_src = g->d_esrc + off;
        a.g.edge_dst = g->d_edst + off;
        a.g.n_edge_local = (uint32_t)(g->blk_off[(size_t)q + 1] - off);
        a.g.n_neg = (uint32_t)g->n_local;
        a.g.neg_shift = a.g.neg_rank = 0;  // negatives are local rows of the context shard
        a.g.vsrc_at = g->d_vsub + g->vsub_off[(size_t)q];
        a.g.n_vsrc = (uint32_t)(g->vsub_off[(size_t)q + 1] - g->vsub_off[(size_t)q]);
        a.Wv = (T*)vslot;
        a.Wc = (T*)m->tab[1];
        a.same_table = 0;
        a.jobs = jobs;
        a.jobs_rem = p->mode == SMORE_MODE_DETERMINISTIC ? 0 : (int)(n_samples % (uint64_t)L.warps);  // nothing lost to rounding
        m->st_samples = 0;
        m->st_ms = 0;
        if (n_samples == 0 || a.g.n_edge_local == 0) return SMORE_OK;
        Timer t;
        if (int rc = t.start()) return rc;
        kern<<<L.blocks, kBlockThreads, smem>>>(a);
        g_launches++;
        CU(cudaGetLastError());
        if (int rc = t.stop(&m->st_ms)) return rc;
        m->st_samples = jobs * (uint64_t)L.warps + (uint64_t)a.jobs_rem;
        return collect_stats(m, L.warps);
    });
}

// MF trainer (src/model/MF.cpp:50-98): LINE's sampling loop, one table in both roles, jobs = total / workers counted from 0.
template <typename T>
int train_mf_t(smore_model_s* m, const smore_train_params* p) {
    return dispatch_dim<T>(m->dim, [&](auto cfg) -> int {
        using C = decltype(cfg);
        void (*kern)(TrainArgs<T>) = k_line<C, false, 0, 1>;
        const size_t smem = batch_smem_bytes<T>(0, p->negative_samples, 0);
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        Launch L;
        if (p->mode == SMORE_MODE_DETERMINISTIC) { L.blocks = 1; L.warps = 1; }
        else if (int rc = pick_grid(kern, smem, effective_max_warps(p, m->rows), p->total, L)) return rc;
        const uint64_t trips = p->total / (uint64_t)L.warps;
        if (int rc = init_state(m, L.warps, 0, p->alpha, p)) return rc;
        TrainArgs<T> a = base_args<T>(m, p, L.warps, (double)p->total, 1, 0, 0);
        a.jobs = trips;
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
// and splits its own share of the samples evenly over them. Two streams: `sc` carries everything that moves rows
// (requests, gather, the three all-to-alls, apply), `su` the update kernels; two buffer sets alternate, so that while
// super-batch i is being updated the rows of super-batch i+1 are already being requested, gathered and shipped:
//     sc:  PREP(0)  PREP(1)  FINISH(0)  PREP(2)  FINISH(1)  ...
//     su:  UPDATE(0)         UPDATE(1)           UPDATE(2)  ...        UPDATE(i) waits for PREP(i), FINISH(i) for UPDATE(i)
// PREP(i+1) gathers rows that UPDATE(i) may still be changing and FINISH(i) has not yet corrected: the staged copies
// are one super-batch staler than without the overlap; the owner still adds exactly `returned - sent`.
namespace {
// The update kernel is persistent: its CTAs occupy every SM they are given for the whole launch, so a communication kernel
// (NCCL send/recv, the gather / apply kernels) launched meanwhile would simply queue behind it and nothing would overlap
// (measured with per-phase events: the all-to-all that returns the rows took 1.6 ms per super-batch, 1.3 ms of it waiting
// for the update kernel of the next super-batch to drain). The update stream therefore lives in a GREEN CONTEXT that owns
// all SMs but `reserve` (SMORE_EXCH_RESERVE_SMS, default 32, 0 = off; the driver rounds partitions to its granularity:
// 32 -> a 120-SM partition on B200), and the remaining SMs are always free for whatever the communication stream launches.
struct ExchStreams {
    cudaStream_t sc = nullptr, su = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr}, updated[2] = {nullptr, nullptr}, all_done = nullptr;
    int su_sms = 0;  // SMs of the update stream's partition (0: the whole device)
    void* gctx = nullptr;
    int carve(int reserve) {
        // driver entry points through the runtime: libcuda is not a link-time dependency of this library
        using FGetRes = CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType);
        using FSplit = CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
        using FDesc = CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int);
        using FCreate = CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
        using FStream = CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int);
        FGetRes get_res = nullptr; FSplit split = nullptr; FDesc gen_desc = nullptr; FCreate create = nullptr; FStream mk_stream = nullptr;
        cudaDriverEntryPointQueryResult q;
        auto sym = [&](const char* name, void** fp) {
            return cudaGetDriverEntryPoint(name, fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fp;
        };
        if (!sym("cuDeviceGetDevResource", (void**)&get_res) || !sym("cuDevSmResourceSplitByCount", (void**)&split) ||
            !sym("cuDevResourceGenerateDesc", (void**)&gen_desc) || !sym("cuGreenCtxCreate", (void**)&create) ||
            !sym("cuGreenCtxStreamCreate", (void**)&mk_stream))
            return -1;
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return -1;
        CUdevResource all, part, rest;
        if (get_res((CUdevice)dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return -1;
        const int total = (int)all.sm.smCount;
        if (reserve <= 0 || reserve >= total) return -1;
        unsigned int groups = 1;
        if (split(&part, &groups, &all, &rest, 0, (unsigned)(total - reserve)) != CUDA_SUCCESS || groups != 1) return -1;
        if ((int)part.sm.smCount >= total) return -1;  // nothing left over for the communication side
        CUdevResourceDesc desc;
        if (gen_desc(&desc, &part, 1) != CUDA_SUCCESS) return -1;
        CUgreenCtx g = nullptr;
        if (create(&g, desc, (CUdevice)dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return -1;
        CUstream st = nullptr;
        if (mk_stream(&st, g, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return -1;
        gctx = g;
        su = (cudaStream_t)st;
        su_sms = (int)part.sm.smCount;
        return 0;
    }
    int init() {
        if (sc) return SMORE_OK;
        CU(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));
        int reserve = 32;  // 2 GPUs, configs[1]: 837 M updates/s without the carve-out, 935 M with 32 SMs set aside
        if (const char* e = getenv("SMORE_EXCH_RESERVE_SMS")) reserve = atoi(e);
        if (reserve <= 0 || carve(reserve) != 0) {
            cudaGetLastError();
            su_sms = 0;
            CU(cudaStreamCreateWithFlags(&su, cudaStreamNonBlocking));
        }
        for (int b = 0; b < 2; ++b) {
            CU(cudaEventCreateWithFlags(&ready[b], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&updated[b], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&all_done, cudaEventDisableTiming));
        if (getenv("SMORE_VERBOSE"))
            fprintf(stderr, "[smore_b200] exchange mode: update stream on %s (reserve request %d)\n",
                    su_sms ? (std::to_string(su_sms) + " SMs of a green context").c_str() : "the whole device (no carve-out)", reserve);
        return SMORE_OK;
    }
};
// One set per device: a process that drives shards on several devices (or switches device between calls) must not launch on
// streams / wait on events that belong to another device.
int exch_streams(ExchStreams** out) {
    static std::mutex mu;
    static std::map<int, ExchStreams> per_dev;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    ExchStreams& xs = per_dev[dev];
    if (int rc = xs.init()) return rc;
    *out = &xs;
    return SMORE_OK;
}
// Timing events of a verbose run: destroyed on every exit path.
struct EventBag {
    std::vector<cudaEvent_t> v;
    cudaEvent_t make() {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        v.push_back(e);
        return e;
    }
    ~EventBag() {
        for (cudaEvent_t e : v) cudaEventDestroy(e);
    }
};
}  // namespace

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
        ExchStreams* xsp = nullptr;
        if (int rc = exch_streams(&xsp)) return rc;
        ExchStreams& g_xs = *xsp;
        cudaStream_t sc = g_xs.sc, su = g_xs.su;
        Launch L;
        if (p->mode == SMORE_MODE_DETERMINISTIC) { L.blocks = 1; L.warps = 1; }
        else if (int rc = pick_grid(kern, smem, effective_max_warps(p, m0->rows * 2), p->total, L, g_xs.su_sms)) return rc;
        const size_t row_bytes = (size_t)m0->dim * sizeof(T);
        const uint64_t wps = (uint64_t)batch_wps(2, p->negative_samples);
        const uint64_t per_sb = (uint64_t)m0->xch->superbatch * (uint64_t)world;
        const uint64_t nsb = std::max<uint64_t>(1, (p->total + per_sb - 1) / per_sb);
        struct Shard {
            TrainArgs<T> a;
            uint64_t trips, jobs_sb, prepared, done;  // per-warp samples: total, per super-batch, requested, launched
            uint64_t jobs_of[2];                      // per-warp samples of the super-batch held by each buffer set
            uint32_t hmask;
        };
        std::vector<Shard> sh((size_t)n);
        // Set-up allocates EVERYTHING the super-batches will need, sized for the worst case (a request list holds unique
        // rows: at most min(samples of the super-batch, rows the owner holds) per requester), so that no rank can run out of
        // memory between two collectives; the ranks then agree on the outcome BEFORE the first exchange (tr.agree), and a
        // rank whose set-up failed makes every rank return instead of leaving its peers inside an all-to-all.
        auto setup = [&](int i) -> int {
            smore_model_s* m = ms[i];
            smore_exchange_s* x = m->xch;
            const uint64_t total_local = (uint64_t)llround((double)p->total * m->g->src_mass_frac);
            const uint64_t jobs = total_local / (uint64_t)L.warps;
            Shard& s = sh[(size_t)i];
            s.trips = cpp ? (jobs > 0 ? jobs - 1 : 0) : jobs;
            s.jobs_sb = (s.trips + nsb - 1) / nsb;
            s.prepared = s.done = 0;
            if (int rc = init_state(m, L.warps, cpp ? 1 : 0, p->alpha, p)) return rc;
            s.a = base_args<T>(m, p, L.warps, (double)total_local, cpp ? 1 : 0, vtab, ctab, m->g->src_mass_frac);
            if (n > 1) s.a.stream_base += (uint64_t)m->g->rank << 20;  // one call drives all the shards: disjoint streams
            const uint64_t sb_samples = std::max<uint64_t>(1, s.jobs_sb * (uint64_t)L.warps);
            uint64_t hsize = 1024;
            while (hsize < 2 * sb_samples) hsize <<= 1;
            if (hsize > (1ull << 30)) return fail(SMORE_E_INVALID, "super-batch too large");
            s.hmask = (uint32_t)(hsize - 1);
            x->req_stride = (int64_t)sb_samples;
            // all ranks together draw total / nsb samples per super-batch (each rounds its share up to a multiple of its warps)
            const uint64_t all_sb = (uint64_t)((double)p->total / (double)nsb) + (uint64_t)L.warps * (uint64_t)world;
            const size_t out_max = (size_t)std::max<uint64_t>(1, std::min<uint64_t>(sb_samples, (uint64_t)(m->g->V - m->rows)));
            const size_t in_max = (size_t)std::max<uint64_t>(1, std::min<uint64_t>(all_sb, (uint64_t)(world - 1) * (uint64_t)m->rows));
            for (int b = 0; b < 2; ++b) {
                ExchSet& e = x->set[b];
                if (int rc = e.hkey.ensure(hsize * 4)) return rc;
                if (int rc = e.hval.ensure(hsize * 4)) return rc;
                if (int rc = e.req.ensure((size_t)world * sb_samples * 4)) return rc;
                if (int rc = e.cnt.ensure(2 * kMaxWorld * 4)) return rc;
                if (int rc = e.off.ensure(kMaxWorld * 4)) return rc;
                if (int rc = e.wrk.ensure(out_max * row_bytes)) return rc;
                if (int rc = e.req_in.ensure(in_max * 4)) return rc;
                if (int rc = e.sent.ensure(in_max * row_bytes)) return rc;
                if (int rc = e.back.ensure(in_max * row_bytes)) return rc;
            }
            if (int rc = x->errors.ensure(sizeof(int))) return rc;
            CU(cudaMemset(x->errors.p, 0, sizeof(int)));
            x->st_rows_moved = 0;
            x->st_superbatches = nsb;
            return SMORE_OK;
        };
        int setup_rc = SMORE_OK;
        for (int i = 0; i < n && !setup_rc; ++i) setup_rc = setup(i);
        if (int rc = tr.agree(setup_rc)) return setup_rc ? setup_rc : rc;
        // SMORE_VERBOSE: device time of the phases on the communication stream, summed over the super-batches
        const bool verbose = getenv("SMORE_VERBOSE") != nullptr;
        enum { P_REQ, P_COUNTS, P_A2A_REQ, P_GATHER, P_ROWS_OUT, P_WAIT_UPD, P_ROWS_BACK, P_APPLY, P_N };
        static const char* const pname[P_N] = {"requests", "counts", "a2a(req)", "gather", "a2a(rows out)", "wait(update)",
                                               "a2a(rows back)", "apply"};
        EventBag pev, uev;  // (destroyed on every exit path)
        std::vector<int> pkind;
        auto mark = [&](int kind) {  // end of phase `kind` on sc (kind < 0: start marker)
            if (!verbose) return;
            if (cudaEvent_t e = pev.make()) {
                cudaEventRecord(e, sc);
                pkind.push_back(kind);
            }
        };
        auto dev_of = [&](int i, int b) {
            smore_exchange_s* x = ms[i]->xch;
            ExchSet& e = x->set[b];
            return ExchDev{(int32_t*)e.hkey.p, (int32_t*)e.hval.p, sh[(size_t)i].hmask, (int32_t*)e.req.p, x->req_stride,
                           (int32_t*)e.cnt.p, (const int32_t*)e.off.p, e.wrk.p, (int*)x->errors.p,
                           x->n_hot ? (const uint32_t*)x->hot.p : nullptr};
        };
        int dev = 0, sms = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int row_blocks = sms * 8;

        // PREP: which remote rows does super-batch `sb` need? -> request lists -> owners gather -> rows to the requesters
        auto prep = [&](uint64_t sb) -> int {
            const int b = (int)(sb & 1);
            mark(-1);
            for (int i = 0; i < n; ++i) {
                Shard& s = sh[(size_t)i];
                ExchSet& e = ms[i]->xch->set[b];
                const uint64_t jobs = std::min<uint64_t>(s.jobs_sb, s.trips - s.prepared);
                s.jobs_of[b] = jobs;
                CU(cudaMemsetAsync(e.hkey.p, 0xff, ((size_t)s.hmask + 1) * 4, sc));
                CU(cudaMemsetAsync(e.cnt.p, 0, 2 * kMaxWorld * 4, sc));
                if (jobs) {
                    k_line_requests<<<L.blocks, kBlockThreads, 0, sc>>>(s.a.g, dev_of(i, b), s.prepared * wps, L.warps, jobs,
                                                                        s.a.seed, s.a.stream_base, s.a.K);
                    g_launches++;
                }
                s.prepared += jobs;
            }
            CU(cudaGetLastError());
            mark(P_REQ);
            if (int rc = tr.counts(ms, n, b, sc)) return rc;
            mark(P_COUNTS);
            for (int i = 0; i < n; ++i) {
                smore_exchange_s* x = ms[i]->xch;
                ExchSet& e = x->set[b];
                const size_t n_out = (size_t)e.off_out[kMaxWorld], n_in = (size_t)e.off_in[kMaxWorld];
                // the buffers were sized for the worst case in set-up: nothing is allocated between two collectives
                if (n_out * row_bytes > e.wrk.cap || n_in * row_bytes > e.sent.cap || n_in * 4 > e.req_in.cap)
                    return fail(SMORE_E_CUDA, "exchange mode: rank %d got %zu request rows in / %zu out, above the set-up bound (internal error)",
                                ms[i]->g->rank, n_in, n_out);
                int32_t off32[kMaxWorld];
                for (int r = 0; r < kMaxWorld; ++r) off32[r] = (int32_t)e.off_out[r];
                CU(cudaMemcpyAsync(e.off.p, off32, sizeof(off32), cudaMemcpyHostToDevice, sc));
                x->st_rows_moved += n_out;
            }
            if (int rc = tr.a2a(ms, n, b, ExchTransport::REQ, row_bytes, sc)) return rc;
            mark(P_A2A_REQ);
            for (int i = 0; i < n; ++i) {
                ExchSet& e = ms[i]->xch->set[b];
                const int64_t n_in = e.off_in[kMaxWorld];
                if (n_in) {
                    k_gather_rows<C><<<row_blocks, kBlockThreads, 0, sc>>>((const T*)ms[i]->tab[vtab], (const int32_t*)e.req_in.p,
                                                                           n_in, (T*)e.sent.p, ms[i]->dim);
                    g_launches++;
                }
            }
            mark(P_GATHER);
            if (int rc = tr.a2a(ms, n, b, ExchTransport::ROWS_OUT, row_bytes, sc)) return rc;
            mark(P_ROWS_OUT);
            CU(cudaEventRecord(g_xs.ready[b], sc));
            return SMORE_OK;
        };
        // UPDATE: every row access is local (or a hot row behind a peer mapping)
        auto update = [&](uint64_t sb) -> int {
            const int b = (int)(sb & 1);
            CU(cudaStreamWaitEvent(su, g_xs.ready[b], 0));
            cudaEvent_t u0 = nullptr, u1 = nullptr;
            if (verbose) {
                u0 = uev.make();
                u1 = uev.make();
                if (u0) cudaEventRecord(u0, su);
            }
            for (int i = 0; i < n; ++i) {
                Shard& s = sh[(size_t)i];
                s.a.jobs = s.jobs_of[b];
                s.a.x = dev_of(i, b);
                if (s.a.jobs) {
                    kern<<<L.blocks, kBlockThreads, smem, su>>>(s.a);
                    g_launches++;
                    s.done += s.a.jobs;
                }
            }
            if (u1) cudaEventRecord(u1, su);
            CU(cudaEventRecord(g_xs.updated[b], su));
            return SMORE_OK;
        };
        // FINISH: rows back to their owners, who add what changed
        auto finish = [&](uint64_t sb) -> int {
            const int b = (int)(sb & 1);
            mark(-1);
            CU(cudaStreamWaitEvent(sc, g_xs.updated[b], 0));
            mark(P_WAIT_UPD);
            if (int rc = tr.a2a(ms, n, b, ExchTransport::ROWS_BACK, row_bytes, sc)) return rc;
            mark(P_ROWS_BACK);
            for (int i = 0; i < n; ++i) {
                ExchSet& e = ms[i]->xch->set[b];
                const int64_t n_in = e.off_in[kMaxWorld];
                if (n_in) {
                    k_apply_delta<C><<<row_blocks, kBlockThreads, 0, sc>>>((T*)ms[i]->tab[vtab], (const int32_t*)e.req_in.p, n_in,
                                                                           (const T*)e.back.p, (const T*)e.sent.p, ms[i]->dim);
                    g_launches++;
                }
            }
            CU(cudaGetLastError());
            mark(P_APPLY);
            return SMORE_OK;
        };

        // (init_state / buffer setup above ran on the legacy stream and were synchronous; sc and su do not synchronise
        // with it, so the timer's closing event is ordered behind both explicitly)
        CU(cudaDeviceSynchronize());
        Timer t;
        if (int rc = t.start()) return rc;
        // A failure past this point is a CUDA / NCCL error (nothing else can fail: see set-up). The transport is torn down
        // (NCCL: ncclCommAbort, so that this rank's queued sends / receives end and its streams drain; peers see the
        // communicator fail instead of waiting for rows that will never come) and both streams are drained before the
        // buffers they use can be freed.
        auto pipeline = [&]() -> int {
            if (int rc = prep(0)) return rc;
            for (uint64_t sb = 0; sb < nsb; ++sb) {
                if (int rc = update(sb)) return rc;
                if (sb + 1 < nsb)
                    if (int rc = prep(sb + 1)) return rc;
                if (int rc = finish(sb)) return rc;
            }
            return SMORE_OK;
        };
        if (int rc = pipeline()) {
            tr.abort();
            cudaStreamSynchronize(su);
            cudaStreamSynchronize(sc);
            cudaGetLastError();
            return rc;
        }
        CU(cudaEventRecord(g_xs.all_done, su));
        CU(cudaStreamWaitEvent(0, g_xs.all_done, 0));
        CU(cudaEventRecord(g_xs.all_done, sc));
        CU(cudaStreamWaitEvent(0, g_xs.all_done, 0));
        double ms_total = 0;
        if (int rc = t.stop(&ms_total)) return rc;
        if (verbose) {
            double sum[P_N] = {};
            for (size_t k = 1; k < pev.v.size() && k < pkind.size(); ++k)
                if (pkind[k] >= 0) {
                    float f = 0;
                    cudaEventElapsedTime(&f, pev.v[k - 1], pev.v[k]);
                    sum[pkind[k]] += f;
                }
            fprintf(stderr, "[smore_b200] exchange rank %d: %llu super-batches in %.2f ms; communication stream:", m0->g->rank,
                    (unsigned long long)nsb, ms_total);
            for (int k = 0; k < P_N; ++k) fprintf(stderr, " %s %.2f", pname[k], sum[k]);
            double upd = 0;
            for (size_t k = 0; k + 1 < uev.v.size(); k += 2) {
                float f = 0;
                cudaEventElapsedTime(&f, uev.v[k], uev.v[k + 1]);
                upd += f;
            }
            fprintf(stderr, " ms; update stream: k_line %.2f ms\n", upd);
        }
        for (int i = 0; i < n; ++i) {
            ms[i]->st_ms = ms_total;
            ms[i]->st_samples = sh[(size_t)i].done * (uint64_t)L.warps;
            if (int rc = collect_stats(ms[i], L.warps)) return rc;
            int errs = 0;
            CU(cudaMemcpy(&errs, ms[i]->xch->errors.p, sizeof(int), cudaMemcpyDeviceToHost));
            if (errs) return fail(SMORE_E_CUDA, "exchange mode: %d samples of rank %d drew a source that was not requested (internal error)", errs, ms[i]->g->rank);
        }
        return SMORE_OK;
    });
}
