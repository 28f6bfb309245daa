// Shared by every translation unit of libsmore_b200.so: handles, error plumbing, launch helpers.
#pragma once
#include "../../include/smore_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "host_graph.h"
#include "batch_kernels.cuh"
#include "exchange_host.h"

using namespace smore;

// defined in smore_b200.cu
extern std::atomic<uint64_t> g_launches;
extern int g_device;
int fail(int code, const char* fmt, ...);
int ensure_device();

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(SMORE_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T>
int dev_alloc_copy(T** d, const T* h, size_t n) {
    *d = nullptr;
    if (n == 0) return SMORE_OK;
    CU(cudaMalloc((void**)d, n * sizeof(T)));
    CU(cudaMemcpy(*d, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return SMORE_OK;
}

struct smore_graph_s {
    int sem = 0, neg_method = 0;
    int64_t V = 0, E = 0, n_lines = 0;
    std::vector<int64_t> row_off;
    std::vector<int32_t> col;
    std::vector<double> w;
    std::vector<std::string> names;
    std::vector<int32_t> field;
    bool has_field = false;
    AliasHost vertex_at, negative_at, ctx_at;
    std::vector<double> out_deg, in_deg;
    // row sharding (one process per GPU): this rank owns vertices v with (v & (world-1)) == rank
    int rank = 0, world = 1, shift = 0;
    int64_t n_local = 0;            // owned vertices
    double src_mass_frac = 1.0;     // share of the global source-sampling mass owned by this rank
    bool neg_global = false;        // sharded: negatives drawn over ALL vertices (context rows then cross NVLink too)
    int64_t n_neg = 0;              // entries of the negative table on the device
    int n_models = 0;               // models ever created on this graph (their row counts freeze the sharding; never decremented:
                                    // a model may outlive its graph handle on the caller's side)
    // device
    int64_t* d_row_off = nullptr;
    int32_t* d_col = nullptr;
    uint2 *d_vat = nullptr, *d_nat = nullptr, *d_cat = nullptr, *d_eat = nullptr;
    int32_t *d_esrc = nullptr, *d_edst = nullptr;
    int64_t n_edge_local = 0;
    double* d_prefix = nullptr;
    int32_t* d_field = nullptr;
    int32_t* d_col_sorted = nullptr;  // node2vec: adjacency slices sorted ascending (built on first use)
    double* d_w = nullptr;            // node2vec: raw edge weights on the device
    double* d_lut64 = nullptr;
    float* d_lut32 = nullptr;
    // rotating shards (smore_graph_set_shard_rotating; rotation.cu): the vertex table is cut into nsub = 2*world
    // sub-parts that travel around the ring of ranks; d_eat / d_esrc / d_edst then hold one edge table PER BLOCK
    // (sub-part q of the sources x this rank's contexts), block q at [blk_off[q], blk_off[q+1])
    bool rotating = false;
    bool synthetic = false;           // generated on the device (device_graph.cu): no host CSR, names or host alias tables
    int nsub = 0;
    int64_t sub_cap = 0;              // rows of a slot buffer: first half of a shard = local rows [0, sub_cap)
    std::vector<int64_t> blk_off;     // nsub + 1
    std::vector<double> blk_mass;     // nsub: share of the GLOBAL edge mass in block (q, this rank)
    std::vector<int64_t> sub_rows;    // nsub: rows of sub-part q (ring index q = 2*home + half)
    uint2* d_vsub = nullptr;          // per sub-part: alias table of the source distribution over its rows (split samples)
    std::vector<int64_t> vsub_off;    // nsub + 1

    GraphDev view() const {
        GraphDev g;
        g.V = V; g.E = E;
        g.row_off = d_row_off; g.col = d_col;
        g.vertex_at = d_vat; g.negative_at = d_nat; g.ctx_at = d_cat;
        g.prefix = d_prefix; g.field = d_field; g.sem = sem;
        g.col_sorted = d_col_sorted; g.w = d_w;
        g.n_neg = (uint32_t)(n_neg ? n_neg : n_local);
        g.neg_shift = neg_global ? 0 : shift; g.neg_rank = neg_global ? 0 : rank;
        g.edge_at = d_eat; g.edge_src = d_esrc; g.edge_dst = d_edst; g.n_edge_local = (uint32_t)n_edge_local;
        g.shard_shift = shift; g.shard_rank = rank;
        g.vsrc_at = d_vat; g.n_vsrc = (uint32_t)V;
        return g;
    }
    ~smore_graph_s() {
        cudaFree(d_row_off); cudaFree(d_col); cudaFree(d_vat); cudaFree(d_nat); cudaFree(d_cat);
        cudaFree(d_prefix); cudaFree(d_field); cudaFree(d_lut64); cudaFree(d_lut32);
        cudaFree(d_eat); cudaFree(d_esrc); cudaFree(d_edst); cudaFree(d_vsub); cudaFree(d_col_sorted); cudaFree(d_w);
    }
};

// Rotating vertex table of one rank (rotation.cu). Three slot buffers of sub_cap rows take the roles
//   T(e) = slot[e % 3]        the sub-part trained in episode e,
//   O(e) = slot[(e + 2) % 3]  the sub-part trained in episode e-1, travelling to the next rank during episode e,
//   I(e) = slot[(e + 1) % 3]  receives the previous rank's O(e); becomes T(e+1).
struct smore_rotation_s {
    void* slot[3] = {nullptr, nullptr, nullptr};
    void* next_slot[3] = {nullptr, nullptr, nullptr};  // the next rank's slots (peer mappings, or same-process pointers)
    bool next_opened[3] = {false, false, false};
    cudaStream_t copy_stream = nullptr;
    int64_t episode = 0;   // the next episode to run
    bool sending = false;  // between send_begin and send_end
    bool trained = false;  // train_line_episode(episode) done
    bool at_home(int nsub) const { return nsub > 0 && episode % nsub == 0 && !sending; }
    void* half(int h) const { return slot[(episode + (h ? 2 : 0)) % 3]; }  // valid at home only
    ~smore_rotation_s() {
        for (int k = 0; k < 3; ++k) {
            if (next_opened[k]) cudaIpcCloseMemHandle(next_slot[k]);
            cudaFree(slot[k]);
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
    }
};

struct smore_model_s {
    smore_graph_t g = nullptr;
    int dim = 0, n_tables = 0, dtype = 0;
    void* tab[2] = {nullptr, nullptr};
    int64_t rows = 0;                          // rows held locally (V when not sharded)
    void* peer[2][kMaxWorld] = {};             // shard bases per rank (peer[t][rank] == tab[t])
    bool peer_opened[2][kMaxWorld] = {};       // CUDA-IPC mappings to close
    void* replica[2] = {nullptr, nullptr};     // optional full-size local read replica of a sharded table
    smore_exchange_s* xch = nullptr;           // bulk-exchange mode of a sharded model (smore_model_enable_exchange)
    smore_rotation_s* rot = nullptr;           // rotating vertex table (smore_model_enable_rotation): tab[0] == nullptr then
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;  // smore_model_{set,get}_rows_f32_async
    cudaEvent_t h2d_event[2] = {nullptr, nullptr};            // last upload / read-back enqueued PER TABLE (cross-stream
    cudaEvent_t d2h_event[2] = {nullptr, nullptr};            // ordering: copies of different tables never wait on each other)
    WarpState* d_state = nullptr;
    int state_cap = 0;
    int32_t* d_keys = nullptr;
    int64_t keys_cap = 0;
    // training position after the last train call (saved in checkpoints: smore_model_progress)
    uint64_t ck_seed = 0, ck_next_stream = 0, ck_sched_total = 0, ck_sched_done = 0;
    // stats of the last train call
    uint64_t st_samples = 0, st_pairs = 0, st_words0 = 0, st_tries = 0;
    double st_ms = 0;
    // CPR / TPR: adjacency of the second graph (over ITS vids) and the third table, aux_V x dim in the model's dtype
    int64_t* d_aux_off = nullptr;
    int32_t* d_aux_col = nullptr;
    int64_t aux_V = 0, aux_E = 0;
    void* aux_tab = nullptr;
    // live progress (smore_progress): host-mapped {done, alpha bits, total, running}; schedule units done before this call
    unsigned long long* live = nullptr;
    unsigned long long live_offset = 0;
    unsigned long long* h_stats = nullptr;  // host-mapped {pairs, tries, words of warp 0, alpha bits}: collect_stats
    size_t elem() const { return dtype == SMORE_F64 ? 8 : 4; }
    ~smore_model_s() {
        for (int t = 0; t < 2; ++t)
            for (int r = 0; r < kMaxWorld; ++r)
                if (peer_opened[t][r]) cudaIpcCloseMemHandle(peer[t][r]);
        cudaFree(tab[0]); cudaFree(tab[1]); cudaFree(d_state); cudaFree(d_keys);
        cudaFree(replica[0]); cudaFree(replica[1]);
        if (h2d_stream) cudaStreamDestroy(h2d_stream);
        if (d2h_stream) cudaStreamDestroy(d2h_stream);
        for (int t = 0; t < 2; ++t) {
            if (h2d_event[t]) cudaEventDestroy(h2d_event[t]);
            if (d2h_event[t]) cudaEventDestroy(d2h_event[t]);
        }
        delete xch;
        delete rot;
        if (live) cudaFreeHost(live);
        if (h_stats) cudaFreeHost(h_stats);
        cudaFree(d_aux_off); cudaFree(d_aux_col); cudaFree(aux_tab);
    }
};

// Rows [first, first+n) of a table as contiguous device ranges: fn(device pointer, first local row, rows) -> rc. One range
// for an ordinary table; a rotating table keeps the two halves of the shard in separate slot buffers and can only be
// addressed while its sub-parts are at home (rotation.cu).
template <class F>
int for_row_ranges(smore_model_s* m, int table, int64_t first, int64_t n, F&& fn) {
    const size_t row_bytes = (size_t)m->dim * m->elem();
    if (!(m->rot && table == 0)) return fn((char*)m->tab[table] + (size_t)first * row_bytes, first, n);
    if (!m->rot->at_home(m->g->nsub))
        return fail(SMORE_E_INVALID, "the rotating vertex table is not at its home position (episode %lld of a %d-episode cycle)",
                    (long long)m->rot->episode, m->g->nsub);
    const int64_t cap = m->g->sub_cap;
    if (first < cap) {
        const int64_t k = std::min(n, cap - first);
        if (int rc = fn((char*)m->rot->half(0) + (size_t)first * row_bytes, first, k)) return rc;
        first += k;
        n -= k;
    }
    if (n > 0) return fn((char*)m->rot->half(1) + (size_t)(first - cap) * row_bytes, first, n);
    return SMORE_OK;
}

namespace {

struct Launch {
    int blocks = 0, warps = 0;
};

// Hogwild concurrency policy. Every resident warp is one lock-free worker; the reference runs at most a few dozen of them
// on tables of any size, this backend runs 3 552 per GPU. fp32 rows take their updates atomically (kernels.cuh,
// kAtomicRows), so concurrent updates of a row are never lost -- but they are all computed from the same slightly stale
// values, which multiplies the effective step of a row by the number of workers that hit it at once. On a table with
// millions of rows that number is ~0; on a toy table of a few hundred rows it is dozens and the linear models (MF) diverge.
// Unless the caller fixes max_warps, the grid is therefore capped at one warp per 8 table rows the trainer touches
// (never below 64): no effect from ~30 k rows up, i.e. on every BASELINE config.
inline int effective_max_warps(const smore_train_params* p, int64_t rows_touched) {
    if (p->max_warps > 0) return p->max_warps;
    return (int)std::min<int64_t>(std::max<int64_t>(64, rows_touched / 8), 1 << 30);
}

template <class K>
int pick_grid(K kernel, size_t smem, int max_warps, uint64_t work_items, Launch& out, int sms_override = 0) {
    int dev = 0, sms = 0, occ = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sms_override > 0) sms = sms_override;  // the kernel runs on an SM partition (green context)
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBlockThreads, smem));
    if (occ < 1) return fail(SMORE_E_CUDA, "kernel does not fit on an SM");
    int64_t warps = (int64_t)sms * occ * kWarpsPerBlock;  // persistent grid: every CTA resident, a multiple of the SM count
    if (max_warps > 0) warps = std::min<int64_t>(warps, max_warps);
    warps = std::min<int64_t>(warps, (int64_t)std::max<uint64_t>(work_items, 1));
    out.warps = (int)warps;
    out.blocks = (int)((warps + kWarpsPerBlock - 1) / kWarpsPerBlock);
    return SMORE_OK;
}

// F is a generic lambda taking a RowCfg tag.
template <typename T, class F>
int dispatch_dim(int dim, F&& f) {
    constexpr int V16 = 16 / (int)sizeof(T);  // elements per 128-bit vector
    if (dim <= 0) return fail(SMORE_E_INVALID, "dim must be positive");
    if (dim % (32 * V16) == 0) {
        int nch = dim / (32 * V16);
        if (nch == 1) return f(RowCfg<T, V16, 1, false>{});
        if (nch == 2) return f(RowCfg<T, V16, 2, false>{});
        if (nch == 4) return f(RowCfg<T, V16, 4, false>{});
    }
    if (V16 == 4 && dim == 64) return f(RowCfg<T, 2, 1, false>{});
    if (dim <= 32) return f(RowCfg<T, 1, 1, true>{});
    if (dim <= 64) return f(RowCfg<T, 1, 2, true>{});
    if (dim <= 128) return f(RowCfg<T, 1, 4, true>{});
    return fail(SMORE_E_UNSUPPORTED, "dim=%d unsupported (<=128, or a multiple of %d up to %d)", dim, 32 * V16, 128 * V16);
}

int ensure_state(smore_model_s* m, int warps) {
    if (warps > m->state_cap) {
        cudaFree(m->d_state);
        m->d_state = nullptr;
        CU(cudaMalloc((void**)&m->d_state, (size_t)warps * sizeof(WarpState)));
        m->state_cap = warps;
    }
    return SMORE_OK;
}

int check_train(smore_model_s* m, const smore_train_params* p, int need_tables, bool shard_ok = false) {
    if (!m || !p) return fail(SMORE_E_INVALID, "null model/params");
    if (m->rot || m->g->rotating) return fail(SMORE_E_UNSUPPORTED, "rotating shards are trained episode by episode: smore_rot_send_begin / smore_train_line_episode / smore_rot_send_end");
    if (m->g->world > 1 && !shard_ok) return fail(SMORE_E_UNSUPPORTED, "this trainer does not run on a row-sharded graph yet (LINE does)");
    if (m->g->world > 1 && !(m->xch && m->xch->n_hot == 0))  // (without hot rows the exchange mode never touches a peer)
        for (int t = 0; t < m->n_tables; ++t)
            for (int r = 0; r < m->g->world; ++r)
                if (!m->peer[t][r]) return fail(SMORE_E_INVALID, "table %d: shard of rank %d not connected (smore_model_open_peers)", t, r);
    if (p->semantics != m->g->sem) return fail(SMORE_E_INVALID, "params.semantics (%d) != graph semantics (%d)", p->semantics, m->g->sem);
    if (m->n_tables < need_tables) return fail(SMORE_E_INVALID, "model has %d tables, this trainer needs %d", m->n_tables, need_tables);
    if (p->mode != SMORE_MODE_DETERMINISTIC && p->mode != SMORE_MODE_HOGWILD) return fail(SMORE_E_INVALID, "bad mode");
    if (!(p->alpha > 0)) return fail(SMORE_E_INVALID, "alpha must be > 0");
    return ensure_device();
}

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    int start() {
        CU(cudaEventCreate(&a));
        CU(cudaEventCreate(&b));
        CU(cudaEventRecord(a, 0));
        return SMORE_OK;
    }
    int stop(double* ms) {
        CU(cudaEventRecord(b, 0));
        CU(cudaEventSynchronize(b));
        float f = 0;
        CU(cudaEventElapsedTime(&f, a, b));
        *ms = f;
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        a = b = nullptr;
        return SMORE_OK;
    }
    ~Timer() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

// The train calls keep clear of the copy engines: a host that pipelines its own row transfers next to a train call
// (smore_model_{set,get}_rows_f32_async) would otherwise see the call's small state copies queue behind a 0.5 GB upload on
// the same engine. The per-warp state is therefore initialised by a kernel, and its statistics are reduced by a kernel into
// four host-mapped words {pairs, tries, stream position of warp 0, alpha bits of warp 0}.
__global__ void k_fill_state(WarpState* st, int n, WarpState proto) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st[i] = proto;
}
__global__ void k_reduce_state(const WarpState* st, int n, unsigned long long* out) {
    __shared__ unsigned long long s_pairs, s_tries;
    if (threadIdx.x == 0) s_pairs = s_tries = 0;
    __syncthreads();
    unsigned long long pairs = 0, tries = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        pairs += st[i].pairs;
        tries += st[i].tries;
    }
    atomicAdd(&s_pairs, pairs);
    atomicAdd(&s_tries, tries);
    __syncthreads();
    if (threadIdx.x == 0) {
        out[0] = s_pairs;
        out[1] = s_tries;
        out[2] = st[0].pos;
        out[3] = (unsigned long long)__double_as_longlong(st[0].alpha);
        __threadfence_system();
    }
}

int collect_stats(smore_model_s* m, int warps) {
    unsigned long long bits;
    if (!m->h_stats && cudaHostAlloc((void**)&m->h_stats, 4 * sizeof(unsigned long long), cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        m->h_stats = nullptr;
    }
    unsigned long long* dev = nullptr;
    if (m->h_stats && cudaHostGetDevicePointer((void**)&dev, m->h_stats, 0) != cudaSuccess) {
        cudaGetLastError();
        dev = nullptr;
    }
    if (dev) {
        k_reduce_state<<<1, 256>>>(m->d_state, warps, dev);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(0));
        m->st_pairs = m->h_stats[0];
        m->st_tries = m->h_stats[1];
        m->st_words0 = m->h_stats[2];
        bits = m->h_stats[3];
    } else {  // no mapped host memory on this platform: read the states back
        std::vector<WarpState> st((size_t)warps);
        CU(cudaMemcpy(st.data(), m->d_state, st.size() * sizeof(WarpState), cudaMemcpyDeviceToHost));
        m->st_pairs = m->st_tries = 0;
        for (auto& s : st) {
            m->st_pairs += s.pairs;
            m->st_tries += s.tries;
        }
        m->st_words0 = st[0].pos;
        memcpy(&bits, &st[0].alpha, sizeof(bits));
    }
    if (m->live) {  // the call is over: what the LAST tick saw -> where the schedule really stands
        __atomic_store_n(&m->live[1], bits, __ATOMIC_RELAXED);
        __atomic_store_n(&m->live[0], m->live_offset + (unsigned long long)m->st_samples, __ATOMIC_RELAXED);
        __atomic_store_n(&m->live[3], 0ull, __ATOMIC_RELEASE);
    }
    return SMORE_OK;
}

// L1-allocating gathers are enabled only when the gathered table (this rank's shard) is far larger than what the L1s can
// hold (148 SMs x <= 228 KB = 33 MB): then only genuinely hot rows stay resident. SMORE_L1_GATHER=0/1 overrides.
constexpr size_t kL1GatherMinTableBytes = 256ull << 20;
int set_l1_gather(const smore_model_s* m) {
    int on = (size_t)m->rows * (size_t)m->dim * m->elem() >= kL1GatherMinTableBytes ? 1 : 0;
    if (const char* e = getenv("SMORE_L1_GATHER")) on = atoi(e) != 0;
    // (c_l1_gather is a per-translation-unit symbol and this function is too: the cache below is per TU and per device)
    static std::mutex mu;  // (train calls on different handles may come from different host threads)
    static int last_on = -1, last_dev = -1;
    int dev = -1;
    CU(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (on != last_on || dev != last_dev) {
        CU(cudaMemcpyToSymbol(c_l1_gather, &on, sizeof(int)));
        last_on = on;
        last_dev = dev;
    }
    return SMORE_OK;
}

template <typename T>
TrainArgs<T> base_args(smore_model_s* m, const smore_train_params* p, int warps, double total, int lag, int vtab, int ctab,
                       double scale = 1.0 /* row-sharded: this rank's share of the schedule units */) {
    TrainArgs<T> a{};
    set_l1_gather(m);
    a.g = m->g->view();
    a.Wv = (T*)m->tab[vtab];
    a.Wc = (T*)m->tab[ctab];
    for (int r = 0; r < kMaxWorld; ++r) {
        a.peer_v[r] = (T*)m->peer[vtab][r];
        a.peer_c[r] = (T*)m->peer[ctab][r];
    }
    a.world_shift = m->g->shift;
    a.world_mask = m->g->world - 1;
    a.dim = m->dim;
    a.same_table = vtab == ctab;
    a.lut = sizeof(T) == 8 ? (const T*)m->g->d_lut64 : (const T*)m->g->d_lut32;
    a.seed = p->seed;
    a.stream_base = p->stream_base;
    // a call may be one chunk of a longer LR schedule (sched_total / sched_offset, in the same units as `total`)
    const double sched_total = p->sched_total ? (double)p->sched_total * scale : total;
    a.sched = Sched{p->alpha, sched_total, (uint64_t)warps, lag, (double)p->sched_offset * scale, nullptr};
    if (!m->live) {
        unsigned long long* fresh = nullptr;
        if (cudaHostAlloc((void**)&fresh, 4 * sizeof(unsigned long long), cudaHostAllocMapped) == cudaSuccess) {
            memset(fresh, 0, 4 * sizeof(unsigned long long));
            __atomic_store_n(&m->live, fresh, __ATOMIC_RELEASE);  // (smore_progress may be polling from another thread)
        } else {
            cudaGetLastError();  // no mapped memory: training runs, smore_progress reads zeros
        }
    }
    if (m->live) {  // {done, alpha bits, total, running}
        double a0 = p->alpha;
        if (p->sched_total && p->sched_offset) a0 = std::max(a0 * 1e-4, a0 * (1.0 - (double)p->sched_offset / (double)p->sched_total));
        unsigned long long bits;
        memcpy(&bits, &a0, sizeof(bits));
        m->live_offset = (unsigned long long)a.sched.offset;
        __atomic_store_n(&m->live[0], m->live_offset, __ATOMIC_RELAXED);
        __atomic_store_n(&m->live[1], bits, __ATOMIC_RELAXED);
        __atomic_store_n(&m->live[2], (unsigned long long)sched_total, __ATOMIC_RELAXED);
        __atomic_store_n(&m->live[3], 1ull, __ATOMIC_RELEASE);
        unsigned long long* dev = nullptr;
        if (cudaHostGetDevicePointer((void**)&dev, m->live, 0) == cudaSuccess) a.sched.live = dev;
        else cudaGetLastError();
    }
    a.state = m->d_state;
    a.n_warps = warps;
    a.K = p->negative_samples;
    a.order = p->order;
    a.lambda = (T)p->lambda;
    a.xi = (T)p->xi;
    a.omega = (T)p->omega;
    a.eta = p->eta;
    a.vred = 1;
    if (const char* e = getenv("SMORE_SHARD_VRED")) a.vred = atoi(e) != 0;  // A/B switch (tools/ab_sharded_quality.py)
    return a;
}

int init_state(smore_model_s* m, int warps, uint64_t count0, double alpha, const smore_train_params* p = nullptr) {
    if (p) {  // where the run stands once this call returns (walk models: sched units are walks, recorded by the trainer)
        m->ck_seed = p->seed;
        m->ck_next_stream = std::max<uint64_t>(m->ck_next_stream, p->stream_base + (uint64_t)warps);
        m->ck_sched_total = p->sched_total ? p->sched_total : p->total;
        m->ck_sched_done = (p->sched_total ? p->sched_offset : 0) + p->total;
    }
    if (p && p->sched_total && p->sched_offset) {  // resume a longer schedule: alpha_t = alpha * max(1e-4, 1 - done/total)
        const double a = alpha * (1.0 - (double)p->sched_offset / (double)p->sched_total);
        alpha = a < alpha * 0.0001 ? alpha * 0.0001 : a;
    }
    const WarpState proto{0, count0, ((uint64_t)kMonitor + (uint64_t)warps - 1) / (uint64_t)warps, alpha, 0, 0};
    if (int rc = ensure_state(m, warps)) return rc;
    k_fill_state<<<(warps + 255) / 256, 256>>>(m->d_state, warps, proto);  // (a kernel, not a copy: see collect_stats)
    CU(cudaGetLastError());
    return SMORE_OK;
}

constexpr size_t smem_rings = (size_t)kWarpsPerBlock * 256 * sizeof(uint32_t);
template <typename T>
constexpr size_t smem_line() { return smem_rings + 1008 * sizeof(T); }
template <typename T>
constexpr size_t smem_walk() { return smem_line<T>() + (size_t)kWarpsPerBlock * kMaxWalkLen * (sizeof(int32_t) + 1); }


}  // namespace

// one translation unit per (trainer, element type): train_*.cu
template <typename T>
int train_line_t(smore_model_s* m, const smore_train_params* p);
// bulk-exchange mode: ms = the shards driven by this process (1 with the NCCL transport, all of them with the local one)
template <typename T>
int train_line_exchange_t(smore_model_s** ms, int n, const smore_train_params* p, ExchTransport& tr);
template <typename T>
int train_walk_t(smore_model_s* m, const smore_train_params* p, int walklets, int n2v = 0);
template <typename T>
int train_hpe_t(smore_model_s* m, const smore_train_params* p);
template <typename T>
int train_mf_t(smore_model_s* m, const smore_train_params* p);
// rotating shards: one block (sub-part q of the vertex table in `vslot` x this rank's contexts), n_samples updates
template <typename T>
int train_line_block_t(smore_model_s* m, const smore_train_params* p, int q, void* vslot, uint64_t n_samples);
template <typename T>
int train_bpr_block_t(smore_model_s* m, const smore_train_params* p, int q, void* vslot, uint64_t n_samples);
template <typename T>
int train_ranking_t(smore_model_s* m, const smore_train_params* p, int kind);
