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

