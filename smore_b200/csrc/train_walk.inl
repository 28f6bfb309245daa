// DeepWalk / Walklets trainer
#include "host_common.h"

template <typename T>
int train_walk_t(smore_model_s* m, const smore_train_params* p, int walklets, int n2v) {
    return dispatch_dim<T>(m->dim, [&](auto cfg) -> int {
        using C = decltype(cfg);
        auto kern = k_walk<C>;
        const size_t smem = smem_walk<T>();
        const bool cpp = p->semantics == SMORE_SEM_CPP;
        const int64_t V = m->g->V;
        Launch L;
        if (p->mode == SMORE_MODE_DETERMINISTIC) { L.blocks = 1; L.warps = 1; }
        else if (int rc = pick_grid(kern, smem, effective_max_warps(p, m->rows * 2), (uint64_t)V, L)) return rc;
        const double total = (double)((unsigned long long)p->walk_times * (unsigned long long)V);
        if (int rc = init_state(m, L.warps, 0, p->alpha, p)) return rc;
        TrainArgs<T> a = base_args<T>(m, p, L.warps, total, 0, 0, 1);
        a.steps = p->walk_steps;
        a.w0 = p->window_min;
        a.w1 = p->window_max;
        a.walklets = walklets;
        a.n2v = n2v;
        if (n2v) {
            a.n2v_pinv = 1.0 / p->n2v_p;  // the reference divides per neighbour (node2vec.go:133,139): the same IEEE quotient
            a.n2v_qinv = 1.0 / p->n2v_q;
            a.g.col_sorted = m->g->d_col_sorted;
            a.g.w = m->g->d_w;
        }
        if (m->keys_cap < V) {
            cudaFree(m->d_keys);
            m->d_keys = nullptr;
            CU(cudaMalloc((void**)&m->d_keys, (size_t)V * sizeof(int32_t)));
            m->keys_cap = V;
        }
        a.keys = m->d_keys;
        std::vector<int32_t> keys((size_t)V);
        HostStream shuffle(p->seed, kShuffleStream);
        int64_t walks_left = p->max_walks >= 0 ? p->max_walks : (int64_t)p->walk_times * V;
        uint64_t done = 0;
        m->st_ms = 0;
        for (int t = 0; t < p->walk_times && walks_left > 0; ++t) {
            // per-epoch Fisher-Yates (DeepWalk.cpp:124-131 with libc rand() := shuffle-stream word >> 1;
            // deepwalk.go:84-92 with rand.Int63n(n) := umulhi32(word, n)). Walklets draws it but walks in id order.
            for (int64_t v = 0; v < V; ++v) keys[(size_t)v] = (int32_t)v;
            for (int64_t v = 0; v < V; ++v) {
                uint32_t k = shuffle.next();
                int64_t j = cpp ? (int64_t)(int)(v + (int64_t)(k >> 1) % (V - v)) : v + (int64_t)(((uint64_t)k * (uint64_t)(V - v)) >> 32);
                std::swap(keys[(size_t)v], keys[(size_t)j]);
            }
            if (walklets)
                for (int64_t v = 0; v < V; ++v) keys[(size_t)v] = (int32_t)v;
            CU(cudaMemcpy(m->d_keys, keys.data(), (size_t)V * sizeof(int32_t), cudaMemcpyHostToDevice));
            a.n_walks = std::min<int64_t>(V, walks_left);
            Timer tm;
            double ms = 0;
            if (int rc = tm.start()) return rc;
            kern<<<L.blocks, kBlockThreads, smem>>>(a);
            g_launches++;
            CU(cudaGetLastError());
            if (int rc = tm.stop(&ms)) return rc;
            m->st_ms += ms;
            walks_left -= a.n_walks;
            done += (uint64_t)a.n_walks;
        }
        m->st_samples = done;
        return collect_stats(m, L.warps);
    });
}


// HPE trainer (src/model/HPE.cpp:93-147): jobs = total / workers, count from 0, LR read before the counter bump.
template <typename T>
int train_hpe_t(smore_model_s* m, const smore_train_params* p) {
    return dispatch_dim<T>(m->dim, [&](auto cfg) -> int {
        using C = decltype(cfg);
        void (*kern)(TrainArgs<T>) = k_hpe<C>;
        size_t smem = smem_line<T>();
        if constexpr (sizeof(T) == 4) {  // fp32 Hogwild: the restructured kernel (kernels.cuh, k_hpe_fast)
            if (p->mode != SMORE_MODE_DETERMINISTIC && p->walk_steps >= 1 && p->walk_steps <= kHpeMaxSteps &&
                (p->walk_steps + 1) * p->negative_samples <= 32 && !getenv("SMORE_HPE_EXACT_STREAM")) {
                kern = k_hpe_fast<C>;
                smem = 1008 * sizeof(T);
            }
        }
        Launch L;
        if (p->mode == SMORE_MODE_DETERMINISTIC) { L.blocks = 1; L.warps = 1; }
        else if (int rc = pick_grid(kern, smem, effective_max_warps(p, m->rows * 2), p->total, L)) return rc;
        const uint64_t trips = p->total / (uint64_t)L.warps;
        if (int rc = init_state(m, L.warps, 0, p->alpha, p)) return rc;
        TrainArgs<T> a = base_args<T>(m, p, L.warps, (double)p->total, 1, 0, 1);
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
