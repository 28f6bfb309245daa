// Rotating shards: the multi-GPU mode of the row-sharded store for graphs whose tables are too large for fine-grained
// peer access (include/smore_b200.h, "rotating shards").
//
// Both tables are row-sharded as in the peer-access mode (vertex v -> rank v & (G-1), local row v >> log2 G). The CONTEXT
// shard of a rank never moves. The VERTEX shard is cut into two halves, so that the whole table consists of 2G sub-parts,
// and the sub-parts travel around the ring of ranks: in episode e rank g trains the block
//       sources in sub-part q = (2g - e) mod 2G   x   contexts owned by g        (negatives: contexts owned by g)
// out of its own HBM only -- the single-GPU kernel, no remote access at all -- while the sub-part it trained in episode
// e-1 is on its way to rank g+1 as ONE contiguous copy-engine transfer over NVLink (cudaMemcpyAsync into a CUDA-IPC
// mapping of the neighbour's slot buffer), hidden behind the update kernel. After 2G episodes (one cycle) every sub-part
// has met every context shard once and is back home. Each vertex row exists exactly once at any time: nothing is stale,
// nothing is lost, and the union of the blocks' samples -- block (q, g) runs total_cycle * mass(q, g) of them -- is exactly
// the global edge distribution of the unsharded samplers.
// This is the block-cyclic schedule of GraphVite / PyTorch-BigGraph; the reference has nothing comparable (one shared
// table, src/model/LINE.cpp:162-191).
#include "host_common.h"

namespace {

int ring_q(int rank, int64_t episode, int nsub) {  // sub-part trained by `rank` in `episode`
    int64_t q = (2ll * rank - episode) % nsub;
    return (int)(q < 0 ? q + nsub : q);
}

}  // namespace

extern "C" {

int smore_graph_set_shard_rotating(smore_graph_t g, int rank, int world) {
    if (!g) return fail(SMORE_E_INVALID, "null graph");
    if (world < 2 || world > kMaxWorld || (world & (world - 1))) return fail(SMORE_E_INVALID, "world must be 2, 4 or 8");
    if (rank < 0 || rank >= world) return fail(SMORE_E_INVALID, "rank out of range");
    if (g->n_models > 0) return fail(SMORE_E_INVALID, "%d model(s) were already created on this graph: shard the graph first", g->n_models);
    if (g->world != 1) return fail(SMORE_E_INVALID, "the graph is already sharded");
    if (int rc = ensure_device()) return rc;
    int shift = 0;
    while ((1 << shift) < world) ++shift;
    const int64_t V = g->V;
    if (V < 4ll * world) return fail(SMORE_E_INVALID, "the graph is too small to be cut into %d sub-parts", 2 * world);
    const int nsub = 2 * world;
    const int64_t nl = (V - rank + world - 1) / world;
    const int64_t shard_cap = (V + world - 1) / world;
    const int64_t sub_cap = (shard_cap + 1) / 2;
    auto local_rows = [&](int r) { return (V - r + world - 1) / world; };
    // edge probabilities of the unsharded samplers, as in smore_graph_set_shard:
    //   C++: out_deg(v1)^0.75 / sum * w^0.75 / sum_{e of v1} w^0.75 ;  Go: out_deg(v1) / sum * w / out_deg(v1)
    const double pw = g->sem == SMORE_SEM_CPP ? 0.75 : 1.0;
    auto wpow = [pw](double x) { return pw == 0.75 ? pow075(x) : x; };
    std::vector<double> psrc((size_t)V), nrm((size_t)V, 0.0);
    parallel_for(V, 1 << 14, [&](int64_t vb, int64_t ve, int) {
        for (int64_t v = vb; v < ve; ++v) {
            psrc[(size_t)v] = g->out_deg[(size_t)v] > 0 ? std::pow(g->out_deg[(size_t)v], pw) : 0.0;
            for (int64_t e = g->row_off[(size_t)v]; e < g->row_off[(size_t)v + 1]; ++e) nrm[(size_t)v] += wpow(g->w[(size_t)e]);
        }
    });
    double src_sum = 0;
    for (int64_t v = 0; v < V; ++v) src_sum += psrc[(size_t)v];
    if (!(src_sum > 0)) return fail(SMORE_E_INVALID, "the graph has no edges");
    auto sub_of = [&](int64_t v, int64_t* row) {  // ring index q = 2*home + half, and the row inside the sub-part
        const int home = (int)(v & (world - 1));
        const int64_t l = v >> shift;
        const int half = l >= sub_cap ? 1 : 0;
        *row = l - (half ? sub_cap : 0);
        return 2 * home + half;
    };
    // pass 1: block sizes and masses; pass 2: fill (sources ascending inside a block: deterministic)
    std::vector<int64_t> cnt((size_t)nsub, 0);
    std::vector<double> mass((size_t)nsub, 0.0);
    double mass_all = 0;
    for (int64_t v = 0; v < V; ++v) {
        if (!(nrm[(size_t)v] > 0)) continue;
        int64_t row;
        const int q = sub_of(v, &row);
        const double pv = psrc[(size_t)v] / src_sum / nrm[(size_t)v];
        for (int64_t e = g->row_off[(size_t)v]; e < g->row_off[(size_t)v + 1]; ++e) {
            const double pr = pv * wpow(g->w[(size_t)e]);
            mass_all += pr;
            if ((g->col[(size_t)e] & (world - 1)) == rank) {
                cnt[(size_t)q]++;
                mass[(size_t)q] += pr;
            }
        }
    }
    std::vector<int64_t> off((size_t)nsub + 1, 0);
    for (int q = 0; q < nsub; ++q) off[(size_t)q + 1] = off[(size_t)q] + cnt[(size_t)q];
    const int64_t ne = off[(size_t)nsub];
    if (ne == 0) return fail(SMORE_E_INVALID, "rank %d owns no edge target", rank);
    if (ne >= (1ll << 32)) return fail(SMORE_E_UNSUPPORTED, "rank %d owns %lld CSR entries; edge tables are indexed with 32-bit draws", rank, (long long)ne);
    std::vector<double> pe((size_t)ne);
    std::vector<int32_t> esrc((size_t)ne), edst((size_t)ne);
    std::vector<int64_t> fill(off.begin(), off.end() - 1);
    for (int64_t v = 0; v < V; ++v) {
        if (!(nrm[(size_t)v] > 0)) continue;
        int64_t row;
        const int q = sub_of(v, &row);
        const double pv = psrc[(size_t)v] / src_sum / nrm[(size_t)v];
        for (int64_t e = g->row_off[(size_t)v]; e < g->row_off[(size_t)v + 1]; ++e)
            if ((g->col[(size_t)e] & (world - 1)) == rank) {
                const int64_t k = fill[(size_t)q]++;
                pe[(size_t)k] = pv * wpow(g->w[(size_t)e]);
                esrc[(size_t)k] = (int32_t)row;
                edst[(size_t)k] = (int32_t)(g->col[(size_t)e] >> shift);
            }
    }
    // one Vose table per block (alias indices relative to the block), built by the host threads block by block
    std::vector<uint2> packed((size_t)ne);
    for (int q = 0; q < nsub; ++q) {
        const int64_t o = off[(size_t)q], n = cnt[(size_t)q];
        if (n == 0) continue;
        AliasHost t = alias_method_go(pe.data() + o, n, 1.0);
        parallel_for(n, 1 << 16, [&](int64_t b, int64_t e, int) {
            for (int64_t i = b; i < e; ++i) {
                PackedAlias pa = pack_alias(t.prob[(size_t)i], t.alias[(size_t)i], (uint32_t)i);
                packed[(size_t)(o + i)] = make_uint2(pa.thr, pa.alias);
            }
        });
    }
    // negatives: the owned vertices, local row ids (same construction as the global table)
    std::vector<double> neg((size_t)nl);
    for (int64_t l = 0; l < nl; ++l) {
        const int64_t v = l * world + rank;
        const double in = g->in_deg[(size_t)v], out = g->out_deg[(size_t)v];
        if (g->sem == SMORE_SEM_GO || g->neg_method == SMORE_NEG_DEGREES) neg[(size_t)l] = in + out;
        else if (g->neg_method == SMORE_NEG_IN_DEGREES) neg[(size_t)l] = in;
        else neg[(size_t)l] = in == 0 ? 0 : 1;
    }
    g->negative_at = g->sem == SMORE_SEM_CPP ? alias_method_cpp(neg.data(), nl) : alias_method_go(neg.data(), nl, 0.75);
    std::vector<uint2> npacked((size_t)nl);
    for (int64_t i = 0; i < nl; ++i) {
        PackedAlias pa = pack_alias(g->negative_at.prob[(size_t)i], g->negative_at.alias[(size_t)i], (uint32_t)i);
        npacked[(size_t)i] = make_uint2(pa.thr, pa.alias);
    }
    // split samples: the second vertex of a sample comes from the source distribution restricted to the resident sub-part
    // (one table per sub-part, ids = rows inside the sub-part)
    std::vector<int64_t> vsub_off((size_t)nsub + 1, 0);
    for (int q = 0; q < nsub; ++q) {
        const int64_t rows = local_rows(q >> 1);
        vsub_off[(size_t)q + 1] = vsub_off[(size_t)q] + ((q & 1) ? std::max<int64_t>(0, rows - sub_cap) : std::min(rows, sub_cap));
    }
    std::vector<uint2> vpacked((size_t)vsub_off[(size_t)nsub]);
    {
        std::vector<double> wsub;
        for (int q = 0; q < nsub; ++q) {
            const int64_t n = vsub_off[(size_t)q + 1] - vsub_off[(size_t)q];
            if (n == 0) continue;
            wsub.resize((size_t)n);
            for (int64_t r = 0; r < n; ++r) wsub[(size_t)r] = psrc[(size_t)(((r + ((q & 1) ? sub_cap : 0)) << shift) + (q >> 1))];
            AliasHost t = alias_method_go(wsub.data(), n, 1.0);
            for (int64_t i = 0; i < n; ++i) {
                PackedAlias pa = pack_alias(t.prob[(size_t)i], t.alias[(size_t)i], (uint32_t)i);
                vpacked[(size_t)(vsub_off[(size_t)q] + i)] = make_uint2(pa.thr, pa.alias);
            }
        }
    }
    cudaFree(g->d_eat); cudaFree(g->d_esrc); cudaFree(g->d_edst); cudaFree(g->d_nat); cudaFree(g->d_vsub);
    g->d_eat = nullptr; g->d_esrc = g->d_edst = nullptr; g->d_nat = nullptr; g->d_vsub = nullptr;
    if (int rc = dev_alloc_copy(&g->d_vsub, vpacked.data(), vpacked.size())) return rc;
    g->vsub_off = vsub_off;
    if (int rc = dev_alloc_copy(&g->d_eat, packed.data(), packed.size())) return rc;
    if (int rc = dev_alloc_copy(&g->d_esrc, esrc.data(), esrc.size())) return rc;
    if (int rc = dev_alloc_copy(&g->d_edst, edst.data(), edst.size())) return rc;
    if (int rc = dev_alloc_copy(&g->d_nat, npacked.data(), npacked.size())) return rc;
    g->n_edge_local = ne;
    g->rank = rank;
    g->world = world;
    g->shift = shift;
    g->n_local = nl;
    g->n_neg = nl;
    g->neg_global = false;
    g->rotating = true;
    g->nsub = nsub;
    g->sub_cap = sub_cap;
    g->blk_off = off;
    g->blk_mass.assign((size_t)nsub, 0.0);
    double own = 0;
    for (int q = 0; q < nsub; ++q) {
        g->blk_mass[(size_t)q] = mass_all > 0 ? mass[(size_t)q] / mass_all : 0.0;
        own += g->blk_mass[(size_t)q];
    }
    g->src_mass_frac = own;
    g->sub_rows.assign((size_t)nsub, 0);
    for (int r = 0; r < world; ++r) {
        const int64_t rows = local_rows(r);
        g->sub_rows[(size_t)(2 * r)] = std::min(rows, sub_cap);
        g->sub_rows[(size_t)(2 * r + 1)] = std::max<int64_t>(0, rows - sub_cap);
    }
    return SMORE_OK;
}

int smore_graph_rotation_info(smore_graph_t g, int* n_sub, int64_t* sub_cap, double* block_mass, int64_t* block_edges) {
    if (!g || !g->rotating) return fail(SMORE_E_INVALID, "the graph is not set up for rotating shards");
    if (n_sub) *n_sub = g->nsub;
    if (sub_cap) *sub_cap = g->sub_cap;
    for (int q = 0; q < g->nsub; ++q) {
        if (block_mass) block_mass[q] = g->blk_mass[(size_t)q];
        if (block_edges) block_edges[q] = g->blk_off[(size_t)q + 1] - g->blk_off[(size_t)q];
    }
    return SMORE_OK;
}

int smore_model_enable_rotation(smore_model_t m) {
    if (!m) return fail(SMORE_E_INVALID, "null model");
    smore_graph_s* g = m->g;
    if (!g->rotating) return fail(SMORE_E_INVALID, "call smore_graph_set_shard_rotating before creating the model");
    if (m->n_tables != 2) return fail(SMORE_E_INVALID, "rotating shards need a vertex and a context table");
    if (m->rot) return SMORE_OK;
    if (int rc = ensure_device()) return rc;
    const size_t row_bytes = (size_t)m->dim * m->elem();
    smore_rotation_s* r = new smore_rotation_s();
    for (int k = 0; k < 3; ++k) {
        cudaError_t e = cudaMalloc(&r->slot[k], (size_t)g->sub_cap * row_bytes);
        if (e != cudaSuccess) {
            delete r;
            return fail(SMORE_E_NOMEM, "cudaMalloc of a %zu-byte slot buffer failed: %s", (size_t)g->sub_cap * row_bytes, cudaGetErrorString(e));
        }
    }
    cudaError_t e = cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete r;
        return fail(SMORE_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    // the rows the table already holds move into the home slots: first half -> T(0) = slot[0], second half -> O(0) = slot[2]
    const int64_t n0 = std::min(m->rows, g->sub_cap), n1 = m->rows - n0;
    e = cudaMemcpy(r->slot[0], m->tab[0], (size_t)n0 * row_bytes, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess && n1 > 0)
        e = cudaMemcpy(r->slot[2], (const char*)m->tab[0] + (size_t)n0 * row_bytes, (size_t)n1 * row_bytes, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) {
        delete r;
        return fail(SMORE_E_CUDA, "moving the vertex rows into the slot buffers: %s", cudaGetErrorString(e));
    }
    cudaFree(m->tab[0]);
    m->tab[0] = nullptr;
    m->peer[0][g->rank] = nullptr;
    m->rot = r;
    return SMORE_OK;
}

int smore_model_rot_ipc_handles(smore_model_t m, void* handles192) {
    if (!m || !m->rot || !handles192) return fail(SMORE_E_INVALID, "rotation not enabled on this model");
    for (int k = 0; k < 3; ++k) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, m->rot->slot[k]));
        memcpy((char*)handles192 + 64 * k, &h, 64);
    }
    return SMORE_OK;
}

int smore_model_rot_open_next(smore_model_t m, const void* handles192) {
    if (!m || !m->rot || !handles192) return fail(SMORE_E_INVALID, "rotation not enabled on this model");
    if (int rc = ensure_device()) return rc;
    for (int k = 0; k < 3; ++k) {
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles192 + 64 * k, 64);
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        m->rot->next_slot[k] = p;
        m->rot->next_opened[k] = true;
    }
    return SMORE_OK;
}

int smore_model_rot_slot_ptrs(smore_model_t m, void** ptrs3) {
    if (!m || !m->rot || !ptrs3) return fail(SMORE_E_INVALID, "rotation not enabled on this model");
    for (int k = 0; k < 3; ++k) ptrs3[k] = m->rot->slot[k];
    return SMORE_OK;
}

int smore_model_rot_set_next_ptrs(smore_model_t m, void* const* ptrs3) {
    if (!m || !m->rot || !ptrs3) return fail(SMORE_E_INVALID, "rotation not enabled on this model");
    for (int k = 0; k < 3; ++k) m->rot->next_slot[k] = ptrs3[k];
    return SMORE_OK;
}

int smore_rot_position(smore_model_t m, int64_t* episode, int* at_home, int* training_subpart) {
    if (!m || !m->rot) return fail(SMORE_E_INVALID, "rotation not enabled on this model");
    if (episode) *episode = m->rot->episode;
    if (at_home) *at_home = m->rot->at_home(m->g->nsub) ? 1 : 0;
    if (training_subpart) *training_subpart = ring_q(m->g->rank, m->rot->episode, m->g->nsub);
    return SMORE_OK;
}

// Episode e, step 1 of 3 (non-blocking): the sub-part trained in episode e-1 starts its trip to the next rank.
int smore_rot_send_begin(smore_model_t m, int64_t episode) {
    if (!m || !m->rot) return fail(SMORE_E_INVALID, "rotation not enabled on this model");
    smore_rotation_s* r = m->rot;
    if (episode != r->episode || r->sending) return fail(SMORE_E_INVALID, "send_begin(%lld): the model is at episode %lld%s", (long long)episode, (long long)r->episode, r->sending ? " (send already in flight)" : "");
    for (int k = 0; k < 3; ++k)
        if (!r->next_slot[k]) return fail(SMORE_E_INVALID, "the next rank's slot buffers are not connected (smore_model_rot_open_next)");
    if (int rc = ensure_device()) return rc;
    const smore_graph_s* g = m->g;
    const int q_out = ring_q(g->rank, episode - 1, g->nsub);  // what O(e) holds
    const size_t bytes = (size_t)g->sub_rows[(size_t)q_out] * (size_t)m->dim * m->elem();
    // (the trainers run on the legacy stream and are synchronous: O(e) is final by the time this call is made)
    if (bytes) CU(cudaMemcpyAsync(r->next_slot[(episode + 1) % 3], r->slot[(episode + 2) % 3], bytes, cudaMemcpyDefault, r->copy_stream));
    r->sending = true;
    r->trained = false;
    return SMORE_OK;
}

// Step 2 of 3: one block of updates out of local HBM. p->total = samples of ALL ranks in this episode; this rank runs
// total * 2G * mass(q, rank) of them (every block's share of a cycle), sched_total / sched_offset are global sample counts.
static int train_episode(smore_model_t m, const smore_train_params* p, int64_t episode, bool bpr) {
    if (!m || !p || !m->rot) return fail(SMORE_E_INVALID, "rotation not enabled on this model");
    smore_rotation_s* r = m->rot;
    if (episode != r->episode || !r->sending || r->trained)
        return fail(SMORE_E_INVALID, "train episode %lld: call order is send_begin, train_*_episode, send_end (model at episode %lld)", (long long)episode, (long long)r->episode);
    if (p->semantics != m->g->sem) return fail(SMORE_E_INVALID, "params.semantics (%d) != graph semantics (%d)", p->semantics, m->g->sem);
    if (p->mode != SMORE_MODE_DETERMINISTIC && p->mode != SMORE_MODE_HOGWILD) return fail(SMORE_E_INVALID, "bad mode");
    if (!(p->alpha > 0)) return fail(SMORE_E_INVALID, "alpha must be > 0");
    if (bpr && p->semantics != SMORE_SEM_GO)
        return fail(SMORE_E_UNSUPPORTED, "rotating shards: the Go BPR (user and item tables) only; the C++ BPR trains ONE table in both roles");
    if (!bpr && p->order != 2) return fail(SMORE_E_UNSUPPORTED, "rotating shards: LINE order 2 only (order 1 shares one table between both roles)");
    if (!bpr && (p->negative_samples < 0 || p->negative_samples > 31)) return fail(SMORE_E_UNSUPPORTED, "negative_samples must be in [0,31]");
    if (int rc = ensure_device()) return rc;
    const smore_graph_s* g = m->g;
    const int q = ring_q(g->rank, episode, g->nsub);
    const uint64_t n = (uint64_t)llround((double)p->total * (double)g->nsub * g->blk_mass[(size_t)q]);
    void* slot = r->slot[episode % 3];
    int rc;
    if (bpr) rc = m->dtype == SMORE_F64 ? train_bpr_block_t<double>(m, p, q, slot, n) : train_bpr_block_t<float>(m, p, q, slot, n);
    else rc = m->dtype == SMORE_F64 ? train_line_block_t<double>(m, p, q, slot, n) : train_line_block_t<float>(m, p, q, slot, n);
    if (rc) return rc;
    r->trained = true;
    return SMORE_OK;
}
int smore_train_line_episode(smore_model_t m, const smore_train_params* p, int64_t episode) { return train_episode(m, p, episode, false); }
int smore_train_bpr_episode(smore_model_t m, const smore_train_params* p, int64_t episode) { return train_episode(m, p, episode, true); }

// Step 3 of 3: wait for this rank's outgoing copy. The HOST then synchronises the ranks (a barrier of its own transport):
// after it, every rank's I(e) holds the sub-part of episode e+1 and every O(e) may be overwritten.
int smore_rot_send_end(smore_model_t m, int64_t episode) {
    if (!m || !m->rot) return fail(SMORE_E_INVALID, "rotation not enabled on this model");
    smore_rotation_s* r = m->rot;
    if (episode != r->episode || !r->sending) return fail(SMORE_E_INVALID, "send_end(%lld): no send in flight (model at episode %lld)", (long long)episode, (long long)r->episode);
    CU(cudaStreamSynchronize(r->copy_stream));
    r->sending = false;
    r->trained = false;
    r->episode = episode + 1;
    return SMORE_OK;
}

}  // extern "C"
