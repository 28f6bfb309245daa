// BPR / WARP / HOP-Rec trainer
#include "host_common.h"

template <typename T>
int train_ranking_t(smore_model_s* m, const smore_train_params* p, int kind) {
    return dispatch_dim<T>(m->dim, [&](auto cfg) -> int {
        using C = decltype(cfg);
        const bool cpp = p->semantics == SMORE_SEM_CPP;
        void (*kern)(TrainArgs<T>) = kind == RANK_WARP ? k_warp<C> : kind == RANK_HOPREC ? k_hoprec<C> : kind == RANK_SKEWOPT ? k_skewopt<C>
                                     : kind == RANK_CPR ? k_cpr<C> : kind == RANK_TPR ? k_tpr<C>
                                     : cpp ? k_bpr_cpp<C> : k_bpr_go<C>;
        bool warp_fast = false;
        if constexpr (sizeof(T) == 4) {  // WARP, fp32 Hogwild: the restructured kernel (ranking_kernels.cuh, k_warp_fast)
            if (kind == RANK_WARP && p->mode != SMORE_MODE_DETERMINISTIC && !getenv("SMORE_WARP_EXACT_STREAM")) {
                kern = k_warp_fast<C>;
                warp_fast = true;
            }
        }
        if constexpr (sizeof(T) == 4) {  // CPR / TPR, fp32 Hogwild: group sampling (ranking_kernels.cuh, FAST)
            if (kind >= RANK_CPR && p->mode != SMORE_MODE_DETERMINISTIC && !getenv("SMORE_AUX_EXACT_STREAM"))
                kern = kind == RANK_CPR ? k_cpr<C, true> : k_tpr<C, true>;
        }
        if constexpr (sizeof(T) == 4) {  // HOP-Rec, fp32 Hogwild: k_hoprec_fast
            if (kind == RANK_HOPREC && p->mode != SMORE_MODE_DETERMINISTIC && p->walk_steps >= 1 && p->walk_steps <= kHopMaxSteps &&
                !getenv("SMORE_HOPREC_EXACT_STREAM")) {
                kern = k_hoprec_fast<C>;
                warp_fast = true;  // (same shared-memory footprint: the sigmoid table only)
            }
        }
        const size_t smem = warp_fast ? 1008 * sizeof(T) : kind == RANK_SKEWOPT ? batch_smem_bytes<T>(0, kSbprRounds)
                            : kind != RANK_BPR ? smem_line<T>() : cpp ? batch_smem_bytes<T>(0, 5) : batch_smem_bytes<T>(1, 1);
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        Launch L;
        if (p->mode == SMORE_MODE_DETERMINISTIC) { L.blocks = 1; L.warps = 1; }
        else if (int rc = pick_grid(kern, smem, effective_max_warps(p, m->rows * ((kind == RANK_BPR && !cpp) || kind >= RANK_CPR ? 2 : 1)), p->total, L)) return rc;
        // BPR.cpp:73-85 / WARP.cpp / HBPR.cpp: jobs = total / workers, count from 0; bpr.go: sample_times*MaxLine trips
        const uint64_t trips = p->total / (uint64_t)L.warps;
        if (int rc = init_state(m, L.warps, 0, p->alpha, p)) return rc;
        const int ctab = ((kind == RANK_BPR && !cpp) || kind >= RANK_CPR) ? 1 : 0;  // the C++ ranking models pass one table for both roles
        TrainArgs<T> a = base_args<T>(m, p, L.warps, (double)p->total, cpp ? 1 : 0, 0, ctab);
        if (kind >= RANK_CPR) {  // second graph + third table (smore_model_attach_aux)
            a.aux_off = m->d_aux_off;
            a.aux_col = m->d_aux_col;
            a.aux_V = m->aux_V;
            a.aux_tab = (T*)m->aux_tab;
            a.item_reg = (T)p->item_reg;
            a.margin = (T)p->margin;
            a.text_w = (T)p->text_weight;
        }
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


// Rotating shards, Go BPR (bpr.go:84-131 restricted to one block): users of the resident vertex sub-part x the items this
// rank owns, negatives among the rank's own vertices. Same structure as train_line_block_t (train_line.inl).
template <typename T>
int train_bpr_block_t(smore_model_s* m, const smore_train_params* p, int q, void* vslot, uint64_t n_samples) {
    return dispatch_dim<T>(m->dim, [&](auto cfg) -> int {
        using C = decltype(cfg);
        void (*kern)(TrainArgs<T>) = k_bpr_go<C>;
        const size_t smem = batch_smem_bytes<T>(2, 1);
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        Launch L;
        if (p->mode == SMORE_MODE_DETERMINISTIC) { L.blocks = 1; L.warps = 1; }
        else if (int rc = pick_grid(kern, smem, effective_max_warps(p, m->g->sub_rows[(size_t)q] + m->rows), n_samples, L)) return rc;
        if (int rc = init_state(m, L.warps, 0, p->alpha, p)) return rc;
        TrainArgs<T> a = base_args<T>(m, p, L.warps, (double)n_samples, 0, 0, 1, p->total ? (double)n_samples / (double)p->total : 1.0);
        const smore_graph_s* g = m->g;
        const int64_t off = g->blk_off[(size_t)q];
        a.g.edge_at = g->d_eat + off;
        a.g.edge_src = g->d_esrc + off;
        a.g.edge_dst = g->d_edst + off;
        a.g.n_edge_local = (uint32_t)(g->blk_off[(size_t)q + 1] - off);
        a.g.n_neg = (uint32_t)g->n_local;
        a.g.neg_shift = a.g.neg_rank = 0;
        a.Wv = (T*)vslot;
        a.Wc = (T*)m->tab[1];
        a.same_table = 0;
        a.jobs = n_samples / (uint64_t)L.warps;
        a.jobs_rem = p->mode == SMORE_MODE_DETERMINISTIC ? 0 : (int)(n_samples % (uint64_t)L.warps);
        m->st_samples = 0;
        m->st_ms = 0;
        if (n_samples == 0 || a.g.n_edge_local == 0) return SMORE_OK;
        Timer t;
        if (int rc = t.start()) return rc;
        kern<<<L.blocks, kBlockThreads, smem>>>(a);
        g_launches++;
        CU(cudaGetLastError());
        if (int rc = t.stop(&m->st_ms)) return rc;
        m->st_samples = a.jobs * (uint64_t)L.warps + (uint64_t)a.jobs_rem;
        return collect_stats(m, L.warps);
    });
}
