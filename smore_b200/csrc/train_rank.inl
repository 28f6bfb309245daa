// BPR / WARP / HOP-Rec trainer
#include "host_common.h"

template <typename T>
int train_ranking_t(smore_model_s* m, const smore_train_params* p, int kind) {
    return dispatch_dim<T>(m->dim, [&](auto cfg) -> int {
        using C = decltype(cfg);
        const bool cpp = p->semantics == SMORE_SEM_CPP;
        void (*kern)(TrainArgs<T>) = kind == RANK_WARP ? k_warp<C> : kind == RANK_HOPREC ? k_hoprec<C> : kind == RANK_SKEWOPT ? k_skewopt<C>
                                     : cpp ? k_bpr_cpp<C> : k_bpr_go<C>;
        const size_t smem = kind == RANK_SKEWOPT ? batch_smem_bytes<T>(0, kSbprRounds)
                            : kind != RANK_BPR ? smem_line<T>() : cpp ? batch_smem_bytes<T>(0, 5) : batch_smem_bytes<T>(1, 1);
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        Launch L;
        if (p->mode == SMORE_MODE_DETERMINISTIC) { L.blocks = 1; L.warps = 1; }
        else if (int rc = pick_grid(kern, smem, effective_max_warps(p, m->rows * (kind == RANK_BPR && !cpp ? 2 : 1)), p->total, L)) return rc;
        // BPR.cpp:73-85 / WARP.cpp / HBPR.cpp: jobs = total / workers, count from 0; bpr.go: sample_times*MaxLine trips
        const uint64_t trips = p->total / (uint64_t)L.warps;
        if (int rc = init_state(m, L.warps, 0, p->alpha, p)) return rc;
        const int ctab = (kind == RANK_BPR && !cpp) ? 1 : 0;  // the C++ ranking models pass one table for both roles
        TrainArgs<T> a = base_args<T>(m, p, L.warps, (double)p->total, cpp ? 1 : 0, 0, ctab);
        a.jobs = trips;
        a.steps = p->walk_steps;
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
