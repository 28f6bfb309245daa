// Batch-sampled trainers: LINE, BPR (Go), BPR (C++).
//
// In these loops a sample consumes a FIXED number of stream words:
//     C++:  SourceSample p,idx | TargetSample p,idx | K x NegativeSample idx,p      = 4 + 2K   (src/proNet.cpp:623-683)
//     Go :  aliasSample idx,p  | TargetSample r     | K x aliasSample idx,p         = 3 + 2K   (alias.go:93-106,
//                                                                                               pronet.go:257-284)
// so sample s of a worker owns words [s*wps, (s+1)*wps) of its stream and the draws of different samples are
// independent of each other and of the embeddings. The warp resolves 32 samples at a time, ONE PER LANE: the Philox
// words of the batch are generated cooperatively into shared memory, every lane walks its own source -> target chain
// and looks up its K negatives (32 dependent lookup chains in flight per warp instead of one), the ids are parked in
// shared memory, and the warp then applies the 32 updates in stream order while prefetching the rows of the sample
// kLinePrefetch ahead into L2. Word accounting is identical to a worker that samples one at a time, which is what the
// DETERMINISTIC mode (one warp) relies on.
// (A source is never a sink -- zero out-weight gets alias probability 0 -- so the Go "skip" branch of line.go:121-124 /
// bpr.go:103-106, which would shorten a sample to 2 words, cannot fire; if it ever did the sample is dropped at fixed
// width.)
//
// Shared memory: sigmoid LUT | per warp: word buffer [32*wps + 8] | ids [32*(K+2)]  (ids row = v1, v2, neg_0..neg_K-1)
#pragma once
#include <type_traits>

#include "kernels.cuh"
#include "ranking_kernels.cuh"

namespace smore {

// batch modes: 0 = C++ draws, 1 = Go draws, 2 = sharded: one edge draw (idx, p) replaces the source + target draws,
// 3 = sharded + split sample: edge draw, K negatives, then (idx, p) for the second vertex (ids row grows by one slot)
__host__ __device__ inline int batch_wps(int go, int K) { return (go == 3 ? 4 : go == 2 ? 2 : go ? 3 : 4) + 2 * K; }
__host__ __device__ inline int batch_idw(int go, int K) { return K + 2 + (go == 3 ? 1 : 0); }
__host__ __device__ inline int batch_wbuf_words(int go, int K) { return ((32 * batch_wps(go, K) + 8 + 3) / 4) * 4; }
constexpr int kStageDepth = 4;  // vertex rows of a row-sharded table staged ahead per warp (cp.async ring)
template <typename T>
inline size_t batch_smem_bytes(int go, int K, int stage_row_elems = 0) {
    size_t base = 1008 * sizeof(T) + (size_t)kWarpsPerBlock * (size_t)(batch_wbuf_words(go, K) + 32 * batch_idw(go, K)) * 4;
    base = (base + 15) & ~(size_t)15;
    return base + (size_t)kWarpsPerBlock * kStageDepth * (size_t)stage_row_elems * sizeof(T);
}

// SMORE_LINE_PIPE (EXPERIMENT, off): the unsharded C++ LINE kernel fetches ALL rows of the next samples -- vertex row + K + 1
// context rows -- with cp.async into a per-warp ring of shared-memory stages, so that the bytes a warp keeps in flight do not
// drop to zero during the dots and the reds and cost no register. Measured (profiles/r2t_line_pipe_ab.txt): one stage at 3
// resident CTAs 382 M updates/s (the staged rows still have to sit in registers for the dots: spills), three stages at 2
// resident CTAs (48 samples in flight per SM instead of 24) 629 M/s, against 784 M/s for plain 128-bit loads into registers;
// 519 vs 675 M/s at the configs[4] per-GPU footprint. The register-resident gather is already at what the memory system
// takes from one SM; the ring only removes warps. Kept compilable for the next look (-DSMORE_LINE_PIPE=1).
#ifndef SMORE_LINE_PIPE
#define SMORE_LINE_PIPE 0
#endif
constexpr int kPipeRows = 7;   // vertex + kCtxChunk context rows
constexpr int kPipeDepth = 3;  // stages per warp: the rows of the next two samples are in flight while one is computed
template <class C>
__host__ __device__ constexpr bool line_pipe_cfg() {  // fp32 rows of <= 512 bytes (dim <= 128): 2 CTAs x 8 warps x 3 stages x 3.5 KB = 168 KB per SM
    return SMORE_LINE_PIPE && sizeof(typename C::T) == 4 && C::EPL <= 4;
}
template <class C>
inline size_t line_pipe_bytes() {
    return line_pipe_cfg<C>() ? (size_t)kWarpsPerBlock * kPipeDepth * kPipeRows * (size_t)(C::EPL * 32) * sizeof(typename C::T) : 0;
}

// resident CTAs per SM the register allocator must allow: 3 while a row costs <= 16 B per lane (fp32 dim <= 128)
#ifndef SMORE_LINE_MINBLOCKS
#define SMORE_LINE_MINBLOCKS 3
#endif
template <class C>
constexpr int batch_min_blocks() {
    return C::EPL * (int)sizeof(typename C::T) <= 16 ? SMORE_LINE_MINBLOCKS : C::EPL * (int)sizeof(typename C::T) <= 32 ? 2 : 1;
}

struct Batch {
    uint32_t* wbuf;
    int* ids;
    int K, wps, idw;
    bool split;
};

template <typename T>
__device__ __forceinline__ Batch batch_init(int go, int K, int wib) {
    Batch b;
    b.K = K;
    b.wps = batch_wps(go, K);
    b.idw = batch_idw(go, K);
    b.split = go == 3;
    const int wcap = batch_wbuf_words(go, K);
    b.wbuf = reinterpret_cast<uint32_t*>(smem_raw + 1008 * sizeof(T)) + (size_t)wib * (wcap + 32 * b.idw);
    b.ids = reinterpret_cast<int*>(b.wbuf + wcap);
    return b;
}

// Stages 1+2: words of samples [st.pos/wps, +nb) -> shared; one sample per lane -> ids. Advances st.pos.
template <bool GO>
__device__ __forceinline__ void batch_sample(const GraphDev& g, const Batch& b, uint64_t seed, uint64_t stream,
                                             WarpState& st, int nb, int lane) {
    const uint64_t first_blk = st.pos >> 2;
    const int nblk = (int)(((st.pos + (uint64_t)(nb * b.wps) + 3) >> 2) - first_blk);
    __syncwarp();
    for (int k = lane; k < nblk; k += 32) {
        U4 r = philox_block(seed, stream, first_blk + (uint64_t)k);
        *reinterpret_cast<uint4*>(b.wbuf + 4 * k) = make_uint4(r.x, r.y, r.z, r.w);
    }
    __syncwarp();
    if (lane < nb) {
        const uint32_t* wd = b.wbuf + ((uint32_t)st.pos & 3u) + lane * b.wps;
        int* my_ids = b.ids + lane * b.idw;
        const uint32_t V32 = g.n_neg;
        const bool sharded = g.edge_at != nullptr;
        const int noff = sharded ? 2 : GO ? 3 : 4;
        // negatives first: independent lookups, issued 8 at a time
        for (int n0 = 0; n0 < b.K; n0 += 8) {
            uint32_t idx[8];
            uint2 e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (n0 + j < b.K) {
                    idx[j] = index_draw(wd[noff + 2 * (n0 + j)], V32);
                    e[j] = __ldg(g.negative_at + idx[j]);
                }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (n0 + j < b.K) my_ids[2 + n0 + j] = (int)g.global_id(wd[noff + 2 * (n0 + j) + 1] < e[j].x ? idx[j] : e[j].y);
        }
        if (!sharded) {
            // then the dependent source -> target chain
            const int v1 = (int)source_sample(g, wd[0], wd[1]);
            int u;
            const int v2 = (int)target_sample(g, v1, wd[2], GO ? 0u : wd[3], u);
            my_ids[0] = v1;
            my_ids[1] = v2;
        } else {
            // one draw over this rank's edge table: (idx, p) -> local edge -> (source anywhere, target owned here)
            const uint32_t le = alias_pick(g.edge_at, index_draw(wd[0], g.n_edge_local), wd[1]);
            my_ids[0] = __ldg(g.edge_src + le);
            my_ids[1] = __ldg(g.edge_dst + le);
            if (b.split)  // the vertex that takes the negatives: drawn from the source distribution, independent of the edge
                my_ids[b.K + 2] = (int)alias_pick(g.vsrc_at, index_draw(wd[2 + 2 * b.K], g.n_vsrc), wd[3 + 2 * b.K]);
        }
    }
    st.pos += (uint64_t)(nb * b.wps);
    __syncwarp();
}

// L2 prefetch of every row of one sample (ids row `pid`: slot 0 lives in the vertex table, the others in the context
// table). Rows of a peer shard are skipped: peer addresses bypass the local L2.
// SMORE_PREFETCH_MODE: 0 = none (default since round 2), 1 = one prefetch.global.L2 per 128 B line (round 1; ncu shows each
// of them as a ONE-sector request), 2 = one cp.async.bulk.prefetch.L2 per row (the whole 512 B row in one instruction),
// 3 = one prefetch.global.L2 per 32 B sector.
// Measured with the atomic-row kernels (profiles/r2h_prefetch_modes.txt): every prefetch flavour LOSES to none -- the block
// kernel at the configs[4] footprint 592 M updates/s with any of them vs 674 M/s without (ncu: 4.9 KB of DRAM reads per
// update instead of the 3.8 KB the rows and sampler entries account for: prefetched sectors are evicted before the demand
// load arrives and fetched twice), BPR C++ 0.87 -> 0.95 of the HBM peak, MF 0.80 -> 0.84, configs[1] LINE +1..4 %. Round 1's
// kernels (whole-row stores) gained 4 % from mode 1; with `red` updates the rows no longer wait in registers for the store,
// the 24 resident warps per SM cover the DRAM latency on their own, and the prefetches only add traffic.
#ifndef SMORE_PREFETCH_MODE
#define SMORE_PREFETCH_MODE 0
#endif
__device__ __forceinline__ void prefetch_row(const void* row, int row_bytes, int lane_in_row, int lanes_per_row) {
#if SMORE_PREFETCH_MODE == 2
    if (lane_in_row == 0) {
        if ((row_bytes & 15) == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row), "r"(row_bytes) : "memory");
        else
            for (int o = 0; o < row_bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)row + o));
    }
#elif SMORE_PREFETCH_MODE == 3
    for (int o = lane_in_row * 32; o < row_bytes; o += lanes_per_row * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)row + o));
#elif SMORE_PREFETCH_MODE == 1
    for (int o = lane_in_row * 128; o < row_bytes; o += lanes_per_row * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)row + o));
#endif
}
// lanes are dealt to the rows of the sample in groups of 4 (8 rows per pass)
template <typename T, class TV, class TC>
__device__ __forceinline__ void prefetch_sample(const TV& tv, const TC& tc, int rank, const int* pid,
                                                int idw, int lane, int vslot2 = -1 /* split samples: slot of the 2nd vertex */) {
    const int row_bytes = tv.dim * (int)sizeof(T);
    for (int r = lane >> 2; r < idw; r += 8) {
        const int id = pid[r];
        const bool vrow = r == 0 || r == vslot2;
        // (a negative id in slot 0 is a staging row of the exchange mode: ExchView; negative elsewhere = no row)
        if (vrow ? (id < 0 || (id & tv.mask) == (rank & tv.mask)) : (id >= 0 && (id & tc.mask) == (rank & tc.mask)))
            prefetch_row(vrow ? tv.row(id) : tc.row(id), row_bytes, lane & 3, 4);
    }
}

template <typename T>
__device__ __forceinline__ void prefetch_local(const T* Wv, const T* Wc, const int* pid, int idw, int dim, int lane) {
    const int row_bytes = dim * (int)sizeof(T);
    for (int r = lane >> 2; r < idw; r += 8) {
        const int id = pid[r];
        if (id >= 0) prefetch_row((r == 0 ? Wv : Wc) + (size_t)id * dim, row_bytes, lane & 3, 4);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Bulk-exchange mode (row-sharded LINE on graphs whose remote working set is too large for fine-grained peer access).
// Per super-batch and rank:   k_line_requests  -> all-to-all(request lists) -> k_gather_rows -> all-to-all(rows)
//                          -> k_line<.., 3>    -> all-to-all(rows back)     -> k_apply_delta
// The sampler is counter-based, so k_line_requests simply re-derives the SOURCE of every sample the update kernel is
// about to draw (same stream words) -- no id list goes through HBM.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t exch_hash(int v) {
    uint32_t h = (uint32_t)v * 0x9E3779B1u;
    return h ^ (h >> 15);
}

// Returns the staging row of remote vertex v1 (filed by k_line_requests in an earlier launch), or -1 if the vertex is
// not in the table -- which would mean the two kernels disagreed about a sample; the probe is bounded so that such a bug
// surfaces as an error code instead of a hung device.
__device__ __forceinline__ int exch_lookup(const ExchDev& x, int v1) {
    uint32_t h = exch_hash(v1) & x.hmask;
    for (uint32_t probes = 0; probes <= x.hmask; ++probes) {
        const int k = __ldg(x.hkey + h);
        if (k == v1) {
            const int hv = __ldg(x.hval + h);
            return __ldg(x.off + (hv >> 28)) + (hv & 0x0fffffff);
        }
        if (k == -1) break;
        h = (h + 1) & x.hmask;
    }
    return -1;
}

// One lane per sample: words [pos0 + s*wps, +2) of the warp's stream -> edge draw -> source vertex; remote sources are
// inserted into the hash, the lane that creates an entry appends the row to the owner's request list. Must be launched
// with the grid, `jobs`, seed and stream_base of the k_line<.., 3> launch it prepares. pos0 = stream position of every
// warp at the start of that launch: a sample consumes exactly wps words, so the host knows it without reading the warp
// states (the previous super-batch may still be running).
static __global__ void __launch_bounds__(kBlockThreads) k_line_requests(GraphDev g, ExchDev x, uint64_t pos0,
                                                                        int n_warps, uint64_t jobs, uint64_t seed,
                                                                        uint64_t stream_base, int K) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (w >= n_warps) return;
    const uint64_t stream = stream_base + (uint64_t)w;
    const uint64_t wps = (uint64_t)batch_wps(2, K);
    const int mask = (1 << g.shard_shift) - 1;
    for (uint64_t s0 = 0; s0 < jobs; s0 += 32) {  // warp-uniform trip count: the appends below are warp-collective
        const uint64_t s = s0 + (uint64_t)lane;
        bool created = false;
        int v1 = 0, o = 0;
        uint32_t h = 0;
        if (s < jobs) {
            const uint64_t p0 = pos0 + s * wps;
            const U4 b0 = philox_block(seed, stream, p0 >> 2);
            const uint32_t sel = (uint32_t)p0 & 3u;
            uint32_t w0 = sel == 0 ? b0.x : sel == 1 ? b0.y : sel == 2 ? b0.z : b0.w;
            uint32_t w1 = sel == 0 ? b0.y : sel == 1 ? b0.z : b0.w;
            if (sel == 3) w1 = philox_block(seed, stream, (p0 >> 2) + 1).x;
            const uint32_t le = alias_pick(g.edge_at, index_draw(w0, g.n_edge_local), w1);
            v1 = __ldg(g.edge_src + le);
            o = v1 & mask;
            h = exch_hash(v1) & x.hmask;
            if (o != g.shard_rank && !x.is_hot(v1)) {
                for (;;) {
                    const int prev = atomicCAS(x.hkey + h, -1, v1);
                    if (prev == -1) {
                        created = true;
                        break;
                    }
                    if (prev == v1) break;
                    h = (h + 1) & x.hmask;
                }
            }
        }
        // the lanes that created an entry append to their owner's request list: one atomicAdd per owner and warp (every
        // creator of the device would otherwise hammer the same `world` counters)
        __syncwarp();
        const unsigned cm = __ballot_sync(kFull, created);
        if (created) {
            const unsigned grp = __match_any_sync(cm, o);
            const int leader = __ffs(grp) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(x.cnt + o, __popc(grp));
            base = __shfl_sync(grp, base, leader);
            const int j = base + __popc(grp & ((1u << lane) - 1u));
            x.req[(int64_t)o * x.req_stride + j] = v1 >> g.shard_shift;
            x.hval[h] = (o << 28) | j;
        }
    }
}

// Owner side: rows[i] = table[req[i]] (one warp per row; i runs over the packed request lists of all sources).
template <class C>
__global__ void __launch_bounds__(kBlockThreads) k_gather_rows(const typename C::T* __restrict__ table,
                                                               const int32_t* __restrict__ req, int64_t n,
                                                               typename C::T* __restrict__ rows, int dim) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); i < n; i += nw) {
        Row<C> r;
        r.load(table + (size_t)__ldg(req + i) * dim, lane, dim);
        r.store(rows + (size_t)i * dim, lane, dim);
    }
}

// Owner side: table[req[i]] += back[i] - sent[i] with red.global.add (two requesters may return the same row).
template <class C>
__global__ void __launch_bounds__(kBlockThreads) k_apply_delta(typename C::T* __restrict__ table,
                                                               const int32_t* __restrict__ req, int64_t n,
                                                               const typename C::T* __restrict__ back,
                                                               const typename C::T* __restrict__ sent, int dim) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); i < n; i += nw) {
        Row<C> a, b;
        a.load(back + (size_t)i * dim, lane, dim);
        b.load(sent + (size_t)i * dim, lane, dim);
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) a.x[e] = Ar<typename C::T>::sub(a.x[e], b.x[e]);
        row_red_add<C>(table + (size_t)__ldg(req + i) * dim, a, lane, dim);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// LINE: LINE::Train (src/model/LINE.cpp:100-195) / LINE.Train (internal/models/line/line.go:73-150)
// ---------------------------------------------------------------------------------------------------------------
// SHARD: 0 = one GPU; 1 = row-sharded, remote vertex rows staged over NVLink; 2 = row-sharded with a local read replica of
// the vertex table: rows are read (and updated) in the replica, the delta is pushed to the owner with red.global.add;
// 3 = row-sharded, bulk exchange: remote vertex rows were gathered into the local staging table a.x.wrk beforehand
// (k_line_requests -> all-to-all -> k_gather_rows -> all-to-all), so every access of this kernel is to local HBM --
// except the rows of HOT vertices (ExchDev::hot), which are reached through the peer mappings.
// KIND: 0 = skip-gram pair update (LINE), 1 = MF: the same sampling loop around UpdateFactorizedPair (MF::Train,
// src/model/MF.cpp:70-92, draws exactly what LINE::Train draws), 2 = LINE with split samples (row-sharded modes only:
// update_pair_split, the K negatives go to a second, independently drawn vertex).
template <class C, bool GO, int SHARD, int KIND>
__host__ __device__ constexpr bool line_pipe() {
    return line_pipe_cfg<C>() && SHARD == 0 && KIND == 0 && !GO;
}
template <class C, bool GO, int SHARD, int KIND = 0>
__global__ void __launch_bounds__(kBlockThreads, line_pipe<C, GO, SHARD, KIND>() ? 2 : batch_min_blocks<C>()) k_line(TrainArgs<typename C::T> a) {
    constexpr bool STAGED = SHARD == 1;
    using T = typename C::T;
    const T* lut = stage_lut<T>(a.lut, reinterpret_cast<T*>(smem_raw));
    using TV = typename std::conditional<SHARD == 1, TableView<T>,
                                         typename std::conditional<SHARD == 3, ExchView<T>, DirectView<T>>::type>::type;
    using TC = typename std::conditional<SHARD == 3, OwnedView<T>,
                                         typename std::conditional<SHARD != 0, TableView<T>, DirectView<T>>::type>::type;
    TV tv;
    TC tc;
    TableView<T> owner_v{};  // SHARD == 2: where the vertex deltas are pushed
    if constexpr (SHARD == 1) stage_views<T>(a, tv, tc);
    else if constexpr (SHARD == 3) {
        TableView<T> pv, pc;
        stage_views<T>(a, pv, pc);
        tv = ExchView<T>{pv.base, reinterpret_cast<T*>(a.x.wrk), a.world_shift, a.world_mask, a.dim};
        tc = OwnedView<T>{a.Wc, a.world_shift, a.dim};
    } else if constexpr (SHARD == 2) {
        stage_views<T>(a, owner_v, tc);
        tv = DirectView<T>{a.replica_v, a.dim};
    } else {
        tv = DirectView<T>{a.Wv, a.dim};
        tc = DirectView<T>{a.Wc, a.dim};
    }
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= a.n_warps) return;
    const int bmode = a.g.edge_at ? (KIND == 2 ? 3 : 2) : (GO ? 1 : 0);
    const Batch b = batch_init<T>(bmode, a.K, wib);
    WarpState st = a.state[w];
    const uint64_t stream = a.stream_base + (uint64_t)w;
    const int nrows = a.K + 1;
    // row-sharded: the vertex row of a sample usually lives on another GPU; its ~2-3 us NVLink load is taken off the
    // critical path by staging the rows of the next kStageDepth-1 samples into a per-warp shared-memory ring
    constexpr bool staged = STAGED;  // == row-sharded table (a.world_mask != 0); a template flag keeps the unsharded
                                     // instantiation free of the staging registers
    constexpr int kRowElems = C::EPL * 32;
    T* vstage = nullptr;
    if constexpr (staged) {
        size_t off = 1008 * sizeof(T) + (size_t)kWarpsPerBlock * (size_t)(batch_wbuf_words(bmode, a.K) + 32 * batch_idw(bmode, a.K)) * 4;
        off = (off + 15) & ~(size_t)15;
        vstage = reinterpret_cast<T*>(smem_raw + off) + (size_t)wib * kStageDepth * kRowElems;
    }
    constexpr bool PIPE = line_pipe<C, GO, SHARD, KIND>();
    T* pstage = nullptr;
    if constexpr (PIPE) {
        size_t off = 1008 * sizeof(T) + (size_t)kWarpsPerBlock * (size_t)(batch_wbuf_words(bmode, a.K) + 32 * batch_idw(bmode, a.K)) * 4;
        off = (off + 15) & ~(size_t)15;
        pstage = reinterpret_cast<T*>(smem_raw + off) + (size_t)wib * kPipeDepth * kPipeRows * kRowElems;
    }
    // (one warp = the DETERMINISTIC mode: every sample must see its predecessor's updates, nothing is fetched ahead)
    const bool pipe = PIPE && nrows <= kCtxChunk && !a.same_table && a.n_warps > 1;
    const uint64_t my_jobs = a.jobs + (w < a.jobs_rem ? 1u : 0u);
    for (uint64_t done = 0; done < my_jobs; done += 32) {
        const int nb = (int)min((uint64_t)32, my_jobs - done);
        batch_sample<GO>(a.g, b, a.seed, stream, st, nb, lane);
        if constexpr (PIPE) {
            if (pipe) {
                // every lane copies (and later reads) only the 16-byte pieces it owns: cp.async.wait_group is all the
                // synchronisation the stage needs
                auto issue = [&](int s) {
                    if (s < nb) {
                        const int* sid = b.ids + s * b.idw;
                        T* stg = pstage + (s % kPipeDepth) * (kPipeRows * kRowElems);
                        if (sid[1] >= 0) {
                            row_stage_async<C>(stg, tv.row(sid[0]), lane, a.dim);
                            for (int r = 0; r < nrows; ++r) row_stage_async<C>(stg + (1 + r) * kRowElems, tc.row(sid[1 + r]), lane, a.dim);
                        }
                    }
                    cp_async_commit();
                };
#pragma unroll
                for (int s = 0; s < kPipeDepth - 1; ++s) issue(s);
                for (int s = 0; s < nb; ++s) {
                    const int* sid = b.ids + s * b.idw;
                    const int v1 = sid[0];
                    const int v2 = sid[1];
                    const int my = lane < nrows ? sid[1 + lane] : (-1 - lane);
                    issue(s + kPipeDepth - 1);  // into the stage sample s - 1 has just left
                    cp_async_wait<kPipeDepth - 1>();
                    const T* stg = pstage + (s % kPipeDepth) * (kPipeRows * kRowElems);
                    Row<C> v, c[kCtxChunk];
                    if (v2 >= 0) {
                        row_from_smem<C>(v, stg, lane, a.dim);
#pragma unroll
                        for (int r = 0; r < kCtxChunk; ++r)
                            if (r < nrows) row_from_smem<C>(c[r], stg + (1 + r) * kRowElems, lane, a.dim);
                    }
                    if (v2 < 0) continue;
                    const T alpha = (T)st.alpha;
                    const unsigned peers = __match_any_sync(kFull, my);
                    if (__any_sync(kFull, lane < nrows && __popc(peers) > 1)) {
                        // a context row repeats inside the sample: the ordered path through memory
                        update_pair_cpp<C, TV, TC, false>(tv, tc, a.dim, false, lut, v1, my, nrows, alpha, lane);
                    } else {
                        using A = Ar<T>;
                        Row<C> back;
                        back.zero();
                        T f[kCtxChunk];
                        dots<C, kCtxChunk>(v, c, nrows, f);
#pragma unroll
                        for (int r = 0; r < kCtxChunk; ++r) {
                            if (r < nrows) {
                                const T g = A::mul(A::sub(r == 0 ? (T)1 : (T)0, fast_sigmoid<T>(lut, f[r])), alpha);
#pragma unroll
                                for (int e = 0; e < C::EPL; ++e) {
                                    back.x[e] = A::madd(back.x[e], g, c[r].x[e]);
                                    c[r].x[e] = A::mul(g, v.x[e]);
                                }
                                row_red_add<C>(tc.row(sid[1 + r]), c[r], lane, a.dim);
                            }
                        }
                        row_red_add<C>(tv.row(v1), back, lane, a.dim);
                    }
                    st.count++;
                    st.pairs++;
                    sched_tick(st, a.sched);
                }
                cp_async_wait<0>();
                continue;
            }
        }
        if constexpr (SHARD == 3) {
            // remote source: k_line_requests filed it in the hash; its row now sits in the staging table
            if (lane < nb) {
                const int v1 = b.ids[lane * b.idw];
                if ((v1 & a.world_mask) != a.g.shard_rank && !a.x.is_hot(v1)) {
                    const int row = exch_lookup(a.x, v1);
                    if (row >= 0) b.ids[lane * b.idw] = -2 - row;
                    else {
                        b.ids[lane * b.idw + 1] = -1;  // drop the sample and tell the host
                        atomicAdd(a.x.errors, 1);
                    }
                }
            }
            __syncwarp();
        }
        if constexpr (staged) {
#pragma unroll
            for (int s = 0; s < kStageDepth - 1; ++s) {
                if (s < nb) row_stage_async<C>(vstage + (s % kStageDepth) * kRowElems, tv.row(b.ids[s * b.idw]), lane, a.dim);
                cp_async_commit();
            }
        }
        for (int s = 0; s < nb; ++s) {
            if (s + kLinePrefetch < nb)
                prefetch_sample<T>(tv, tc, a.g.shard_rank, b.ids + (s + kLinePrefetch) * b.idw, b.idw, lane, KIND == 2 ? a.K + 2 : -1);
            const int* sid = b.ids + s * b.idw;
            const int v1 = sid[0];
            const int v2 = sid[1];
            const int my = lane < nrows ? sid[1 + lane] : (-1 - lane);  // lane 0: positive context, lane 1+n: negative n
            Row<C> vrow;
            if constexpr (staged) {
                const int s2 = s + kStageDepth - 1;
                if (s2 < nb) row_stage_async<C>(vstage + (s2 % kStageDepth) * kRowElems, tv.row(b.ids[s2 * b.idw]), lane, a.dim);
                cp_async_commit();
                cp_async_wait<kStageDepth - 1>();
                row_from_smem<C>(vrow, vstage + (s % kStageDepth) * kRowElems, lane, a.dim);
            }
            if (v2 < 0) continue;
            const T alpha = (T)st.alpha;
            T* vpush = nullptr;
            if constexpr (SHARD == 2) vpush = owner_v.row(v1);
            // replica rows never alias shard rows in memory; nor do staging rows (v1 < 0 then never equals a context id)
            const bool same = SHARD == 2 ? false : a.same_table != 0;
            if constexpr (KIND == 1) update_factorized_pair<C, TV, TC>(tv, tc, a.dim, same, v1, my, nrows, alpha, a.lambda, lane);
            else if constexpr (KIND == 2) update_pair_split<C, TV, TC>(tv, tc, a.dim, lut, v1, sid[a.K + 2], my, nrows, alpha, lane, staged ? &vrow : nullptr, SHARD == 1 && a.vred != 0);
            else if (!GO) update_pair_cpp<C, TV, TC, SHARD != 0>(tv, tc, a.dim, same, lut, v1, my, nrows, alpha, lane, staged ? &vrow : nullptr, vpush, SHARD == 1 && a.vred != 0);
            else update_pair_go<C, TV, TC, SHARD != 0>(tv, tc, a.dim, same, a.order == 1, lut, v1, my, nrows, alpha, lane, staged ? &vrow : nullptr, vpush, SHARD == 1 && a.vred != 0);
            st.count++;
            st.pairs++;
            sched_tick(st, a.sched);
        }
        if constexpr (staged) cp_async_wait<0>();
    }
    if (lane == 0) a.state[w] = st;
}

// ---------------------------------------------------------------------------------------------------------------
// Go BPR: BPR.Train (internal/models/bpr/bpr.go:84-131) + UpdateBPRPair (pkg/pronet/optimizer.go:87-117).
// 5 words per sample (K = 1); ids row = user, pos, neg. Users live in Wv, items in Wc.
// ---------------------------------------------------------------------------------------------------------------
#ifndef SMORE_BPR_GROUP
#define SMORE_BPR_GROUP 1
#endif
template <class C>
__global__ void __launch_bounds__(kBlockThreads, batch_min_blocks<C>()) k_bpr_go(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    const T* lut = stage_lut<T>(a.lut, reinterpret_cast<T*>(smem_raw));
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= a.n_warps) return;
    // (rotating shards: the block's edge table replaces the source -> target draws, as in k_line)
    const Batch b = batch_init<T>(a.g.edge_at ? 2 : 1, 1, wib);
    WarpState st = a.state[w];
    const uint64_t stream = a.stream_base + (uint64_t)w;
    const int dim = a.dim;
    const uint64_t my_jobs = a.jobs + (w < a.jobs_rem ? 1u : 0u);
    for (uint64_t done = 0; done < my_jobs; done += 32) {
        const int nb = (int)min((uint64_t)32, my_jobs - done);
        batch_sample<true>(a.g, b, a.seed, stream, st, nb, lane);
        if constexpr (kAtomicRows<C>) {
            // fp32 tables: every row takes its delta with red.global.add (kernels.cuh, kAtomicRows). On a Zipf catalogue the
            // few hottest ITEM rows bound the kernel: each of the 3 552 warps sends its own `red` to the same 512 bytes, and
            // same-address atomics serialise in the L2 slice (measured: throughput x share of the top item = ~100 M row
            // updates/s on Zipf(1.0), 1.8 G samples/s on a flat catalogue). The 32 samples of a batch are therefore grouped
            // by positive item (__match_any_sync): the item row is gathered ONCE per group, carried in registers from one
            // member to the next (each member sees its predecessors' updates, as a sequential worker would) and written back
            // with ONE `red` of the accumulated delta. Samples of different items keep their stream order.
            using A = Ar<T>;
            int my_user = 0, my_pos = -1 - lane, my_neg = 0;
            if (lane < nb) {
                const int* sid = b.ids + lane * b.idw;
                my_user = sid[0];
                if (sid[1] >= 0) my_pos = sid[1];
                my_neg = sid[2];
            }
            unsigned todo = __ballot_sync(kFull, my_pos >= 0);
            // (one warp = the DETERMINISTIC mode: the samples keep the reference's order one by one)
            const unsigned my_grp = (SMORE_BPR_GROUP && a.n_warps > 1) ? __match_any_sync(kFull, my_pos) : 1u << lane;
            while (todo) {
                const int leader = __ffs(todo) - 1;
                unsigned grp = __shfl_sync(kFull, my_grp, leader) & todo;
                todo &= ~grp;
                T* pp = a.Wc + (size_t)__shfl_sync(kFull, my_pos, leader) * dim;
                Row<C> p, dp;
                p.load_ca(pp, lane, dim);
                dp.zero();
                while (grp) {
                    const int s = __ffs(grp) - 1;
                    grp &= grp - 1;
                    T* pv = a.Wv + (size_t)__shfl_sync(kFull, my_user, s) * dim;
                    T* pn = a.Wc + (size_t)__shfl_sync(kFull, my_neg, s) * dim;
                    Row<C> v, n;
                    v.load_ca(pv, lane, dim);
                    n.load_ca(pn, lane, dim);
                    pin(v);
                    pin(n);
                    const T alpha = (T)st.alpha;
                    const T la = A::mul(a.lambda, alpha);
                    Row<C> pn2[2] = {p, n};
                    T sc[2];
                    dots<C, 2>(v, pn2, 2, sc);  // posScore, negScore (optimizer.go:95-100)
                    const T gc = A::mul(alpha, fast_sigmoid<T>(lut, A::sub(sc[1], sc[0])));
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) {
                        const T ve = v.x[e], pe = p.x[e], ne = n.x[e];
                        v.x[e] = A::msub(A::mul(gc, A::sub(pe, ne)), la, ve);  // grad - (lambda*alpha)*w
                        const T d = A::msub(A::mul(gc, ve), la, pe);
                        dp.x[e] = A::add(dp.x[e], d);
                        p.x[e] = A::add(pe, d);
                        n.x[e] = A::msub(A::mul(-gc, ve), la, ne);
                    }
                    row_red_add<C>(pv, v, lane, dim);
                    row_red_add<C>(pn, n, lane, dim);
                    st.count++;
                    st.pairs++;
                    sched_tick(st, a.sched);
                }
                row_red_add<C>(pp, dp, lane, dim);
            }
            continue;
        }
        for (int s = 0; s < nb; ++s) {
            if (s + kLinePrefetch < nb) prefetch_local<T>(a.Wv, a.Wc, b.ids + (s + kLinePrefetch) * b.idw, b.idw, dim, lane);
            const int* sid = b.ids + s * b.idw;
            const int user = sid[0], pos = sid[1], neg = sid[2];
            if (pos < 0) continue;
            const T alpha = (T)st.alpha;
            using A = Ar<T>;
            const T la = A::mul(a.lambda, alpha);  // lambda*alpha*w evaluates left to right
            T* pv = a.Wv + (size_t)user * dim;
            T* pp = a.Wc + (size_t)pos * dim;
            T* pn = a.Wc + (size_t)neg * dim;
            const bool same = pos == neg;
            const bool valias = a.same_table && (user == pos || user == neg);
            Row<C> v, p, n;
            v.load_ca(pv, lane, dim);
            p.load_ca(pp, lane, dim);
            n.load_ca(pn, lane, dim);
            pin(v);
            pin(p);
            pin(n);
            Row<C> pn2[2] = {p, n};
            T sc[2];
            dots<C, 2>(v, pn2, 2, sc);  // posScore, negScore (optimizer.go:95-100)
            const T gc = A::mul(alpha, fast_sigmoid<T>(lut, A::sub(sc[1], sc[0])));
            if (!valias) {
#pragma unroll
                for (int e = 0; e < C::EPL; ++e) {
                    const T vg = A::mul(gc, A::sub(p.x[e], n.x[e]));
                    const T pg = A::mul(gc, v.x[e]);
                    const T ng = A::mul(-gc, v.x[e]);
                    v.x[e] = A::add(v.x[e], A::msub(vg, la, v.x[e]));  // w += grad - (lambda*alpha)*w
                    p.x[e] = A::add(p.x[e], A::msub(pg, la, p.x[e]));
                    const T ncur = same ? p.x[e] : n.x[e];  // pos == neg: the second write lands on the updated row
                    n.x[e] = A::add(ncur, A::msub(ng, la, ncur));
                }
                v.store(pv, lane, dim);
                if (!same) p.store(pp, lane, dim);
                n.store(pn, lane, dim);
            } else {
                // one shared table and the user row coincides with an item row: replay through memory
                for_owned<C>(lane, dim, [&](int, int idx) {
                    const T vg = A::mul(gc, A::sub(ldv(pp + idx), ldv(pn + idx)));
                    const T pg = A::mul(gc, ldv(pv + idx));
                    const T ng = A::mul(-gc, ldv(pv + idx));
                    stv(pv + idx, A::add(ldv(pv + idx), A::msub(vg, la, ldv(pv + idx))));
                    stv(pp + idx, A::add(ldv(pp + idx), A::msub(pg, la, ldv(pp + idx))));
                    stv(pn + idx, A::add(ldv(pn + idx), A::msub(ng, la, ldv(pn + idx))));
                });
            }
            st.count++;
            st.pairs++;
            sched_tick(st, a.sched);
        }
    }
    if (lane == 0) a.state[w] = st;
}

// ---------------------------------------------------------------------------------------------------------------
// C++ BPR: BPR::Train (src/model/BPR.cpp:85-103) + UpdateBPRPair (src/proNet.cpp:1406-1455): 5 rounds, the first
// with the caller's negative. 14 words per sample (K = 5); ids row = user, item, j0..j4; one shared table.
// ---------------------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(kBlockThreads, batch_min_blocks<C>()) k_bpr_cpp(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    const T* lut = stage_lut<T>(a.lut, reinterpret_cast<T*>(smem_raw));
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= a.n_warps) return;
    const Batch b = batch_init<T>(0, 5, wib);
    WarpState st = a.state[w];
    const uint64_t stream = a.stream_base + (uint64_t)w;
    const int dim = a.dim;
    T* W = a.Wv;
    for (uint64_t done = 0; done < a.jobs; done += 32) {
        const int nb = (int)min((uint64_t)32, a.jobs - done);
        batch_sample<false>(a.g, b, a.seed, stream, st, nb, lane);
        for (int s = 0; s < nb; ++s) {
            if (s + kLinePrefetch < nb) prefetch_local<T>(W, W, b.ids + (s + kLinePrefetch) * b.idw, b.idw, dim, lane);
            const int* sid = b.ids + s * b.idw;
            const int v1 = sid[0], v2 = sid[1];
            if (v2 < 0) continue;
            const int my = lane < 7 ? sid[lane] : (-1 - lane);
            const unsigned peers = __match_any_sync(kFull, my);
            const bool dup = __any_sync(kFull, lane < 7 && __popc(peers) > 1);
            st.tries += dup ? 1u : 0u;  // stats: samples that took the ORDERED path (reported as mean_tries)
            using A = Ar<T>;
            const T alpha = (T)st.alpha;
            const T c = A::mul(alpha, (T)0.0025);  // alpha*0.0025*w evaluates left to right
            const T cv = A::mul(alpha, (T)0.025);
            T* pv = W + (size_t)v1 * dim;
            T* pi = W + (size_t)v2 * dim;
            if constexpr (kAtomicRows<C>) {
                // fp32 tables: every row takes its delta with red.global.add; the item row's running value is tracked in
                // registers across the five rounds (coinciding rows just receive several deltas)
                Row<C> v, ri, rj[5], verr, di;
                v.load_ca(pv, lane, dim);
                ri.load_ca(pi, lane, dim);
#pragma unroll
                for (int n = 0; n < 5; ++n) rj[n].load_ca(W + (size_t)sid[2 + n] * dim, lane, dim);
                pin(v);
                pin(ri);
#pragma unroll
                for (int n = 0; n < 5; ++n) pin(rj[n]);  // all seven gathers in flight before the first round
                verr.zero();
                di.zero();
#pragma unroll
                for (int n = 0; n < 5; ++n) {
                    Row<C> cvec;
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) cvec.x[e] = A::sub(ri.x[e], rj[n].x[e]);
                    const T f = dot(v, cvec);
                    const T gg = A::mul(fast_sigmoid<T>(lut, A::sub((T)0, f)), alpha);
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) {
                        verr.x[e] = A::madd(verr.x[e], gg, cvec.x[e]);
                        const T cerr = A::mul(gg, v.x[e]);
                        const T d = A::msub(cerr, c, ri.x[e]);  // item row: -c*w + cerr
                        di.x[e] = A::add(di.x[e], d);
                        ri.x[e] = A::add(ri.x[e], d);
                        rj[n].x[e] = A::sub(A::mul(-c, rj[n].x[e]), cerr);  // negative row: -c*w - cerr
                    }
                    row_red_add<C>(W + (size_t)sid[2 + n] * dim, rj[n], lane, dim);
                }
                row_red_add<C>(pi, di, lane, dim);
#pragma unroll
                for (int e = 0; e < C::EPL; ++e) v.x[e] = A::msub(verr.x[e], cv, v.x[e]);
                row_red_add<C>(pv, v, lane, dim);
            } else if (!dup) {
                Row<C> v, ri, rj[5], verr;
                v.load_ca(pv, lane, dim);
                ri.load_ca(pi, lane, dim);
#pragma unroll
                for (int n = 0; n < 5; ++n) rj[n].load_ca(W + (size_t)sid[2 + n] * dim, lane, dim);
                pin(v);
                pin(ri);
#pragma unroll
                for (int n = 0; n < 5; ++n) pin(rj[n]);  // all seven gathers in flight before the first round
                verr.zero();
#pragma unroll
                for (int n = 0; n < 5; ++n) {
                    Row<C> cvec;
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) cvec.x[e] = A::sub(ri.x[e], rj[n].x[e]);
                    const T f = dot(v, cvec);
                    const T gg = A::mul(fast_sigmoid<T>(lut, A::sub((T)0, f)), alpha);
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) {
                        verr.x[e] = A::madd(verr.x[e], gg, cvec.x[e]);
                        const T cerr = A::mul(gg, v.x[e]);
                        ri.x[e] = A::msub(ri.x[e], c, ri.x[e]);
                        rj[n].x[e] = A::msub(rj[n].x[e], c, rj[n].x[e]);
                        ri.x[e] = A::add(ri.x[e], cerr);
                        rj[n].x[e] = A::sub(rj[n].x[e], cerr);
                    }
                    rj[n].store(W + (size_t)sid[2 + n] * dim, lane, dim);
                }
#pragma unroll
                for (int e = 0; e < C::EPL; ++e) {
                    v.x[e] = A::msub(v.x[e], cv, v.x[e]);
                    v.x[e] = A::add(v.x[e], verr.x[e]);
                }
                ri.store(pi, lane, dim);
                v.store(pv, lane, dim);
            } else {
                Row<C> verr;
                verr.zero();
                for (int n = 0; n < 5; ++n)
                    ordered_round<C>(pv, pi, W + (size_t)sid[2 + n] * dim, dim, lane, lut, alpha, false, (T)0, verr);
                for_owned<C>(lane, dim, [&](int e, int idx) {
                    stv(pv + idx, A::msub(ldv(pv + idx), cv, ldv(pv + idx)));
                    stv(pv + idx, A::add(ldv(pv + idx), verr.x[e]));
                });
            }
            st.count++;
            st.pairs += 5;
            sched_tick(st, a.sched);
        }
    }
    if (lane == 0) a.state[w] = st;
}

// ---------------------------------------------------------------------------------------------------------------
// Skew-OPT: SPR::Train (src/model/SkewOPT.cpp) + UpdateSBPRPair (src/proNet.cpp:1517-1566): 16 rounds, every negative
// drawn inside the update. 36 words per sample (K = 16); ids row = user, item, j0..j15; one shared table. First version:
// every round goes through memory in the reference's order (sbpr_round), the user row moves once at the end by the
// averaged error.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSbprRounds = 16;
template <class C>
__global__ void __launch_bounds__(kBlockThreads) k_skewopt(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    using A = Ar<T>;
    const T* lut = stage_lut<T>(a.lut, reinterpret_cast<T*>(smem_raw));
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= a.n_warps) return;
    const Batch b = batch_init<T>(0, kSbprRounds, wib);
    WarpState st = a.state[w];
    const uint64_t stream = a.stream_base + (uint64_t)w;
    const int dim = a.dim;
    T* W = a.Wv;
    for (uint64_t done = 0; done < a.jobs; done += 32) {
        const int nb = (int)min((uint64_t)32, a.jobs - done);
        batch_sample<false>(a.g, b, a.seed, stream, st, nb, lane);
        for (int s = 0; s < nb; ++s) {
            if (s + kLinePrefetch < nb) prefetch_local<T>(W, W, b.ids + (s + kLinePrefetch) * b.idw, b.idw, dim, lane);
            const int* sid = b.ids + s * b.idw;
            const int v1 = sid[0], v2 = sid[1];
            if (v2 < 0) continue;
            const T alpha = (T)st.alpha;
            T* pv = W + (size_t)v1 * dim;
            T* pi = W + (size_t)v2 * dim;
            Row<C> verr;
            verr.zero();
            int update = 0;
            if constexpr (kAtomicRows<C>) {
                // fp32 tables, FAST path: user and item rows gathered once, the 16 negatives four at a time, every row takes
                // its delta with red.global.add; the item row's running value is tracked in registers across the rounds
                const T c = A::mul(alpha, (T)0.01);
                Row<C> v, ri, di;
                v.load(pv, lane, dim);
                ri.load(pi, lane, dim);
                di.zero();
                for (int n0 = 0; n0 < kSbprRounds; n0 += 4) {
                    Row<C> rj[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) rj[r].load(W + (size_t)sid[2 + n0 + r] * dim, lane, dim);
#pragma unroll
                    for (int r = 0; r < 4; ++r) pin(rj[r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        Row<C> cvec;
#pragma unroll
                        for (int e = 0; e < C::EPL; ++e) cvec.x[e] = A::sub(ri.x[e], rj[r].x[e]);
                        const T f = dot(v, cvec);
                        T g = A::div(A::sub(f, a.xi), a.omega);  // Opt_SBPRSGD (src/proNet.cpp:1070-1098)
                        if (g > (T)2) continue;
                        if (g < (T)-2) g = (T)-2;
                        T g_in = (T)1;
                        for (int i = 0; i < a.eta; ++i) g_in = A::mul(g_in, g);
                        const T chain = A::div(g_in, g);
                        g = A::mul(A::div(A::mul(fast_sigmoid<T>(lut, A::mul((T)-1, g_in)), chain), a.omega), alpha);
                        if (!(g == g)) continue;  // g == 0 exactly: the reference's 0/0; no update instead of a NaN row
                        ++update;
#pragma unroll
                        for (int e = 0; e < C::EPL; ++e) {
                            verr.x[e] = A::madd(verr.x[e], g, cvec.x[e]);
                            const T cerr = A::mul(g, v.x[e]);
                            const T d = A::msub(cerr, c, ri.x[e]);
                            di.x[e] = A::add(di.x[e], d);
                            ri.x[e] = A::add(ri.x[e], d);
                            rj[r].x[e] = A::sub(A::mul(-c, rj[r].x[e]), cerr);
                        }
                        row_red_add<C>(W + (size_t)sid[2 + n0 + r] * dim, rj[r], lane, dim);
                    }
                }
                if (update != 0) {
                    row_red_add<C>(pi, di, lane, dim);
                    const T inv = A::div((T)1, (T)update);
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) v.x[e] = A::msub(A::mul(verr.x[e], inv), c, v.x[e]);
                    row_red_add<C>(pv, v, lane, dim);
                }
            } else {
            for (int n = 0; n < kSbprRounds; ++n)
                if (sbpr_round<C>(pv, pi, W + (size_t)sid[2 + n] * dim, dim, lane, lut, alpha, a.xi, a.omega, a.eta, verr)) ++update;
            if (update != 0) {
                const T c = A::mul(alpha, (T)0.01);
                const T up = (T)update;
                for_owned<C>(lane, dim, [&](int e, int idx) {
                    stv(pv + idx, A::msub(ldv(pv + idx), c, ldv(pv + idx)));
                    stv(pv + idx, A::add(ldv(pv + idx), A::div(verr.x[e], up)));
                });
            }
            }
            st.count++;
            st.pairs += kSbprRounds;
            st.tries += (uint64_t)update;  // stats: accepted rounds
            sched_tick(st, a.sched);
        }
    }
    if (lane == 0) a.state[w] = st;
}

}  // namespace smore
