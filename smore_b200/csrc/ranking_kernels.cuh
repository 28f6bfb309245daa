// Pairwise-ranking kernels with data-dependent draw counts: WARP and HOP-Rec (warp-per-worker over a DrawRing), plus
// the helpers the ranking updates share. (BPR has a fixed draw count per sample and lives in batch_kernels.cuh.)
//
// Every variant first resolves the draws of a sample (they depend on the alias tables / field array only, never on the
// embeddings), then gathers every row it will touch in one batch (FAST path, rows distinct) or replays the reference's
// in-place order through memory (ORDERED path, some rows coincide).
#pragma once
#include "kernels.cuh"

namespace smore {

enum RankKind { RANK_BPR = 0, RANK_WARP = 1, RANK_HOPREC = 2, RANK_SKEWOPT = 3, RANK_CPR = 4, RANK_TPR = 5 };

template <typename T>
__device__ __forceinline__ T ldv(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
template <typename T>
__device__ __forceinline__ void stv(T* p, T v) { *reinterpret_cast<volatile T*>(p) = v; }

// Calls f(e, idx) for every element this lane owns: e = register slot, idx = element index inside the row.
template <class C, class F>
__device__ __forceinline__ void for_owned(int lane, int dim, F&& f) {
#pragma unroll
    for (int c = 0; c < C::NCH; ++c) {
#pragma unroll
        for (int j = 0; j < C::VEC; ++j) {
            int idx = (c * 32 + lane) * C::VEC + j;
            if (!C::MASKED || idx < dim) f(c * C::VEC + j, idx);
        }
    }
}

// Lane-parallel rejection sampling: attempt t uses ring words [t*wpa, (t+1)*wpa). `draw(w0, w1)` proposes a vertex,
// `ok(v)` accepts. Consumes exactly the words the sequential reference loop would. wpa in {1, 2}.
template <class Draw, class Ok>
__device__ __forceinline__ int64_t reject_sample(DrawRing& ring, int lane, uint32_t wpa, Draw&& draw, Ok&& ok) {
    for (;;) {
        ring.ensure();
        uint32_t w0 = ring.peek((uint32_t)lane * wpa);
        uint32_t w1 = ring.peek((uint32_t)lane * wpa + (wpa - 1u));
        int64_t cand = draw(w0, w1);
        unsigned acc = __ballot_sync(kFull, ok(cand));
        if (acc) {
            int first = __ffs(acc) - 1;
            ring.advance((uint32_t)(first + 1) * wpa);
            return __shfl_sync(kFull, cand, first);
        }
        ring.advance(32u * wpa);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One BPR-style round of the C++ ranking updates, ORDERED flavour (rows re-read from memory, per-element writes in the
// reference's order: src/proNet.cpp:1426-1440 / :1485-1504). Returns false if the round was margin-gated away
// (Opt_FBPRSGD, :1023). verr accumulates in registers (it is a local vector in the reference too).
// ---------------------------------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ bool ordered_round(typename C::T* pv, typename C::T* pi, typename C::T* pj, int dim, int lane,
                                              const typename C::T* lut, typename C::T alpha, bool gated,
                                              typename C::T margin, Row<C>& verr) {
    using T = typename C::T;
    using A = Ar<T>;
    Row<C> v, ri, rj, cvec;
    v.load_ca(pv, lane, dim);
    ri.load_ca(pi, lane, dim);
    rj.load_ca(pj, lane, dim);
#pragma unroll
    for (int e = 0; e < C::EPL; ++e) cvec.x[e] = A::sub(ri.x[e], rj.x[e]);
    const T f = dot(v, cvec);
    if (gated && f > margin) return false;
    const T g = A::mul(fast_sigmoid<T>(lut, A::sub((T)0, f)), alpha);
    const T c = A::mul(alpha, (T)0.0025);
    for_owned<C>(lane, dim, [&](int e, int idx) {
        verr.x[e] = A::madd(verr.x[e], g, cvec.x[e]);
        const T cerr = A::mul(g, v.x[e]);
        stv(pi + idx, A::msub(ldv(pi + idx), c, ldv(pi + idx)));
        stv(pj + idx, A::msub(ldv(pj + idx), c, ldv(pj + idx)));
        stv(pi + idx, A::add(ldv(pi + idx), cerr));
        stv(pj + idx, A::sub(ldv(pj + idx), cerr));
    });
    return true;
}

// One round of proNet::UpdateSBPRPair (src/proNet.cpp:1531-1556) with Opt_SBPRSGD (:1070-1098), through memory in the
// reference's order (the rows of one shared table may coincide): g = (f - xi) / omega is gated at g > 2 and clamped at
// -2, the gradient is sigmoid(-g^eta) * g^(eta-1) / omega * alpha. Returns false when the round was gated away.
template <class C>
__device__ __forceinline__ bool sbpr_round(typename C::T* pv, typename C::T* pi, typename C::T* pj, int dim, int lane,
                                           const typename C::T* lut, typename C::T alpha, typename C::T xi,
                                           typename C::T omega, int eta, Row<C>& verr) {
    using T = typename C::T;
    using A = Ar<T>;
    Row<C> v, ri, rj, cvec;
    v.load(pv, lane, dim);
    ri.load(pi, lane, dim);
    rj.load(pj, lane, dim);
#pragma unroll
    for (int e = 0; e < C::EPL; ++e) cvec.x[e] = A::sub(ri.x[e], rj.x[e]);
    const T f = dot(v, cvec);
    T g = A::div(A::sub(f, xi), omega);
    if (g > (T)2) return false;
    if (g < (T)-2) g = (T)-2;
    T g_in = (T)1;
    for (int i = 0; i < eta; ++i) g_in = A::mul(g_in, g);
    const T chain = A::div(g_in, g);
    g = A::div(A::mul(fast_sigmoid<T>(lut, A::mul((T)-1, g_in)), chain), omega);
    g = A::mul(g, alpha);
    const T c = A::mul(alpha, (T)0.01);  // alpha*0.01*w evaluates left to right
    for_owned<C>(lane, dim, [&](int e, int idx) {
        verr.x[e] = A::madd(verr.x[e], g, cvec.x[e]);
        const T cerr = A::mul(g, v.x[e]);
        stv(pi + idx, A::msub(ldv(pi + idx), c, ldv(pi + idx)));
        stv(pj + idx, A::msub(ldv(pj + idx), c, ldv(pj + idx)));
        stv(pi + idx, A::add(ldv(pi + idx), cerr));
        stv(pj + idx, A::sub(ldv(pj + idx), cerr));
    });
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// WARP: WARP::Train (src/model/WARP.cpp:85-103) + UpdateWARPPair (src/proNet.cpp:1353-1403): scan up to 32
// negatives, the first one with margin f < 1 triggers a single BPR step on three rows. Words: source (p, idx),
// target (p, idx), negative (idx, p), then (idx, p) per further negative actually scanned.
// Candidates are evaluated in batches of kWarpBatch rows gathered together; rows do not change before the (single)
// update, so speculative gathers beyond the first violator are exact, only wasted.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWarpBatch = 4;

template <class C>
__global__ void __launch_bounds__(kBlockThreads) k_warp(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    using A = Ar<T>;
    uint32_t* rings = reinterpret_cast<uint32_t*>(smem_raw);
    T* lut_s = reinterpret_cast<T*>(smem_raw + kWarpsPerBlock * 256 * sizeof(uint32_t));
    const T* lut = stage_lut<T>(a.lut, lut_s);
    int lane = threadIdx.x & 31;
    int wib = threadIdx.x >> 5;
    int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= a.n_warps) return;
    WarpState st = a.state[w];
    DrawRing ring;
    ring.init(rings + wib * 256, a.seed, a.stream_base + (uint64_t)w, st.pos, lane);
    const GraphDev& g = a.g;
    const int dim = a.dim;
    T* W = a.Wv;
    for (uint64_t it = 0; it < a.jobs; ++it) {
        ring.ensure();
        int64_t v1 = -1, v2 = -1;
        if (lane == 0) {
            v1 = (int64_t)source_sample(g, ring.peek(0), ring.peek(1));
            int u;
            v2 = target_sample(g, v1, ring.peek(2), ring.peek(3), u);
        }
        v1 = __shfl_sync(kFull, v1, 0);
        v2 = __shfl_sync(kFull, v2, 0);
        if (v2 < 0) {
            ring.advance(6u);
            continue;
        }
        const T alpha = (T)st.alpha;
        T* pv = W + v1 * dim;
        T* pi = W + v2 * dim;
        Row<C> v, ri;
        v.load_ca(pv, lane, dim);
        ri.load_ca(pi, lane, dim);
        int scanned = 0;  // negatives evaluated so far
        bool hit = false;
        // candidate n lives at ring words 4+2n, 5+2n
        for (int base = 0; base < 32 && !hit; base += (base == 0 ? 1 : kWarpBatch)) {
            const int nb = base == 0 ? 1 : min(kWarpBatch, 32 - base);
            int64_t cand = -1;
            if (lane < nb) cand = (int64_t)negative_sample(g, ring.peek(4u + 2u * (uint32_t)(base + lane)), ring.peek(5u + 2u * (uint32_t)(base + lane)));
            Row<C> cvec[kWarpBatch];  // context_vec = w[i] - w[j] (proNet.cpp:1373)
            int64_t jid[kWarpBatch];
            T f[kWarpBatch];
#pragma unroll
            for (int r = 0; r < kWarpBatch; ++r) {
                jid[r] = __shfl_sync(kFull, cand, r);
                if (r < nb) {
                    cvec[r].load_ca(W + jid[r] * dim, lane, dim);
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) cvec[r].x[e] = A::sub(ri.x[e], cvec[r].x[e]);
                }
            }
            dots<C, kWarpBatch>(v, cvec, nb, f);
#pragma unroll
            for (int r = 0; r < kWarpBatch; ++r) {
                if (r < nb && !hit) {
                    ++scanned;
                    if (f[r] < (T)1) {
                        hit = true;
                        const int64_t j = jid[r];
                        T* pj = W + j * dim;
                        const T c = A::mul(alpha, (T)0.0025);
                        // Opt_BPRSGD recomputes f from the same (unchanged) rows, so f[r] is its value (:1380 -> :1059)
                        const T gg = A::mul(fast_sigmoid<T>(lut, A::sub((T)0, f[r])), alpha);
                        if constexpr (kAtomicRows<C>) {
                            // fp32 tables: the three rows take their deltas with red.global.add; the negative's row is
                            // item - difference, no reload
#pragma unroll
                            for (int e = 0; e < C::EPL; ++e) {
                                const T verr = A::mul(gg, cvec[r].x[e]);
                                const T cerr = A::mul(gg, v.x[e]);
                                const T rje = A::sub(ri.x[e], cvec[r].x[e]);
                                cvec[r].x[e] = A::sub(A::mul(-c, rje), cerr);  // negative row: -c*w - cerr
                                ri.x[e] = A::msub(cerr, c, ri.x[e]);           // item row:     -c*w + cerr
                                v.x[e] = A::msub(verr, c, v.x[e]);             // user row:     -c*w + verr
                            }
                            row_red_add<C>(pi, ri, lane, dim);
                            row_red_add<C>(pj, cvec[r], lane, dim);
                            row_red_add<C>(pv, v, lane, dim);
                        } else if (j != v2 && j != v1 && v1 != v2) {
                            Row<C> rj;
                            rj.load_ca(pj, lane, dim);  // only the difference was kept; the row is an L2 hit now
#pragma unroll
                            for (int e = 0; e < C::EPL; ++e) {
                                const T verr = A::mul(gg, cvec[r].x[e]);
                                const T cerr = A::mul(gg, v.x[e]);
                                ri.x[e] = A::msub(ri.x[e], c, ri.x[e]);
                                rj.x[e] = A::msub(rj.x[e], c, rj.x[e]);
                                v.x[e] = A::msub(v.x[e], c, v.x[e]);
                                ri.x[e] = A::add(ri.x[e], cerr);
                                rj.x[e] = A::sub(rj.x[e], cerr);
                                v.x[e] = A::add(v.x[e], verr);
                            }
                            ri.store(pi, lane, dim);
                            rj.store(pj, lane, dim);
                            v.store(pv, lane, dim);
                        } else {
                            // coinciding rows: the reference's per-element order through memory (proNet.cpp:1382-1391)
                            for_owned<C>(lane, dim, [&](int e, int idx) {
                                const T verr = A::mul(gg, cvec[r].x[e]);
                                const T cerr = A::mul(gg, v.x[e]);
                                stv(pi + idx, A::msub(ldv(pi + idx), c, ldv(pi + idx)));
                                stv(pj + idx, A::msub(ldv(pj + idx), c, ldv(pj + idx)));
                                stv(pv + idx, A::msub(ldv(pv + idx), c, ldv(pv + idx)));
                                stv(pi + idx, A::add(ldv(pi + idx), cerr));
                                stv(pj + idx, A::sub(ldv(pj + idx), cerr));
                                stv(pv + idx, A::add(ldv(pv + idx), verr));
                            });
                        }
                    }
                }
            }
        }
        ring.advance(4u + 2u * (uint32_t)scanned);
        st.count++;
        st.pairs++;
        st.tries += (uint64_t)scanned;
        sched_tick(st, a.sched);
    }
    st.pos = ring.pos;
    if (lane == 0) a.state[w] = st;
}

// WARP, fp32 Hogwild throughput path. k_warp above follows the reference's draw order word for word, which makes one
// sample a chain of ~13 dependent memory round trips (alias entry -> CSR record -> context alias -> neighbour id -> two
// rows -> per batch of 4 candidates: alias entries -> rows): 0.61 of the HBM roofline. Hogwild runs are compared with the
// reference statistically, not word for word, so this kernel gives every sample a FIXED 68-word slice of its warp's stream
// (source 2, target 2, 32 candidates x 2; nothing lives in a shared-memory ring, every lane computes the Philox block it
// needs) and reorganises the chain:
//   * the (user, item) pairs of 8 consecutive samples are drawn by 8 lanes at once;
//   * all 32 candidate negatives of a sample are looked up at once, one lane each, together with the user and item rows;
//   * candidate rows are gathered four at a time, the next four already in flight while the current four are scored.
// Same update as k_warp's atomic path: the first candidate with f < 1 triggers one BPR step on three rows (red.global.add).
constexpr int kWarpGroup = 8;
constexpr int kWarpWords = 68;  // words of its stream a sample owns (a multiple of 4: slices start on Philox blocks)

template <class C>
__global__ void __launch_bounds__(kBlockThreads, walk_min_blocks<C>()) k_warp_fast(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    using A = Ar<T>;
    const T* lut = stage_lut<T>(a.lut, reinterpret_cast<T*>(smem_raw));
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (w >= a.n_warps) return;
    WarpState st = a.state[w];
    const GraphDev& g = a.g;
    const int dim = a.dim;
    T* W = a.Wv;
    const uint64_t stream = a.stream_base + (uint64_t)w;
    const uint64_t s_base = (st.pos + kWarpWords - 1) / kWarpWords;  // samples this stream has already served
    for (uint64_t it0 = 0; it0 < a.jobs; it0 += kWarpGroup) {
        const int ng = (int)min((uint64_t)kWarpGroup, a.jobs - it0);
        int64_t mu = -1, mi = -1;
        if (lane < ng) {
            const U4 r = philox_block(a.seed, stream, (uint64_t)(kWarpWords / 4) * (s_base + it0 + (uint64_t)lane));
            mu = (int64_t)source_sample(g, r.x, r.y);
            int used;
            mi = target_sample(g, mu, r.z, r.w, used);
        }
        for (int k = 0; k < ng; ++k) {
            const int64_t v1 = __shfl_sync(kFull, mu, k);
            const int64_t v2 = __shfl_sync(kFull, mi, k);
            if (v2 < 0) continue;
            T* pv = W + v1 * dim;
            T* pi = W + v2 * dim;
            Row<C> v, ri;
            v.load_ca(pv, lane, dim);
            ri.load_ca(pi, lane, dim);
            // candidate `lane`: words 4 + 2*lane, 5 + 2*lane of the sample's slice
            const U4 r = philox_block(a.seed, stream, (uint64_t)(kWarpWords / 4) * (s_base + it0 + (uint64_t)k) + 1u + (uint64_t)(lane >> 1));
            const int64_t cand = (int64_t)negative_sample(g, (lane & 1) ? r.z : r.x, (lane & 1) ? r.w : r.y);
            const T alpha = (T)st.alpha;
            Row<C> bufA[kWarpBatch], bufB[kWarpBatch];
            // candidates base .. base + 3 (candidate 32 does not exist: its slot re-reads candidate 31 and is never scored)
            auto gather = [&](Row<C>(&rows)[kWarpBatch], int base) {
#pragma unroll
                for (int q = 0; q < kWarpBatch; ++q)
                    rows[q].load_ca(W + __shfl_sync(kFull, cand, min(base + q, 31)) * dim, lane, dim);
            };
            int scanned = 0;
            bool hit = false;
            auto step = [&](Row<C>& rj, T f, int idx) {  // rj holds item - candidate; f = user . (item - candidate)
                if (hit || idx >= 32) return;
                ++scanned;
                if (f < (T)1) {
                    hit = true;
                    T* pj = W + __shfl_sync(kFull, cand, idx) * dim;
                    const T c = A::mul(alpha, (T)0.0025);
                    const T gg = A::mul(fast_sigmoid<T>(lut, A::sub((T)0, f)), alpha);
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) {
                        const T verr = A::mul(gg, rj.x[e]);
                        const T cerr = A::mul(gg, v.x[e]);
                        const T rje = A::sub(ri.x[e], rj.x[e]);
                        rj.x[e] = A::sub(A::mul(-c, rje), cerr);
                        ri.x[e] = A::msub(cerr, c, ri.x[e]);
                        v.x[e] = A::msub(verr, c, v.x[e]);
                    }
                    row_red_add<C>(pi, ri, lane, dim);
                    row_red_add<C>(pj, rj, lane, dim);
                    row_red_add<C>(pv, v, lane, dim);
                }
            };
            auto score = [&](Row<C>(&rows)[kWarpBatch], int base) {
                T f[kWarpBatch];
#pragma unroll
                for (int q = 0; q < kWarpBatch; ++q) {
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) rows[q].x[e] = A::sub(ri.x[e], rows[q].x[e]);  // item - candidate
                }
                dots<C, kWarpBatch>(v, rows, kWarpBatch, f);
#pragma unroll
                for (int q = 0; q < kWarpBatch; ++q) step(rows[q], f[q], base + q);
            };
            // the caller's negative (candidate 0) alone first, gathered together with the user and item rows: early in
            // training nearly every sample stops here, and speculative gathers would double its traffic
            {
                Row<C> r0;
                r0.load_ca(W + __shfl_sync(kFull, cand, 0) * dim, lane, dim);
#pragma unroll
                for (int e = 0; e < C::EPL; ++e) r0.x[e] = A::sub(ri.x[e], r0.x[e]);
                step(r0, dot<C>(v, r0), 0);
            }
            if (!hit) {
                gather(bufA, 1);
                for (int base = 1; base < 32 && !hit; base += 2 * kWarpBatch) {
                    if (base + kWarpBatch < 32) gather(bufB, base + kWarpBatch);  // in flight while bufA is scored
                    score(bufA, base);
                    if (hit || base + kWarpBatch >= 32) break;
                    if (base + 2 * kWarpBatch < 32) gather(bufA, base + 2 * kWarpBatch);
                    score(bufB, base + kWarpBatch);
                }
            }
            st.count++;
            st.pairs++;
            st.tries += (uint64_t)scanned;
            sched_tick(st, a.sched);
        }
    }
    st.pos = (s_base + a.jobs) * kWarpWords;
    if (lane == 0) a.state[w] = st;
}

// ---------------------------------------------------------------------------------------------------------------
// HOP-Rec: HBPR::Train (src/model/HBPR.cpp:93-126) + UpdateFBPRPair (src/proNet.cpp:1458-1515).
// Per sampled user: for hop w = 1..walk_steps: item at hop w, field-matched negative, 5 margin-gated BPR rounds
// (rounds > 0 draw UNIFORM negatives with field rejection), learning rate alpha/w, margin 1/w.
// ---------------------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(kBlockThreads) k_hoprec(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    using A = Ar<T>;
    uint32_t* rings = reinterpret_cast<uint32_t*>(smem_raw);
    T* lut_s = reinterpret_cast<T*>(smem_raw + kWarpsPerBlock * 256 * sizeof(uint32_t));
    const T* lut = stage_lut<T>(a.lut, lut_s);
    int lane = threadIdx.x & 31;
    int wib = threadIdx.x >> 5;
    int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= a.n_warps) return;
    WarpState st = a.state[w];
    DrawRing ring;
    ring.init(rings + wib * 256, a.seed, a.stream_base + (uint64_t)w, st.pos, lane);
    const GraphDev& g = a.g;
    const int dim = a.dim;
    T* W = a.Wv;
    const uint32_t V32 = (uint32_t)g.V;
    for (uint64_t it = 0; it < a.jobs; ++it) {
        // user: SourceSample until field == 0 (HBPR.cpp:97-99)
        int64_t vid = reject_sample(ring, lane, 2u,
                                    [&](uint32_t w0, uint32_t w1) { return (int64_t)source_sample(g, w0, w1); },
                                    [&](int64_t c) { return __ldg(g.field + c) == 0; });
        int64_t cid = -1;
        bool dead = false;
        for (int hop = 1; hop <= a.steps && !dead; ++hop) {
            // item at this hop: 1 TargetSample for hop 1, two more for every further hop (HBPR.cpp:100,105-109)
            ring.ensure();
            int64_t c = cid;
            uint32_t used = 0;
            if (lane == 0) {
                int u;
                if (hop == 1) {
                    c = target_sample(g, vid, ring.peek(0), ring.peek(1), u);
                    used = 2;
                } else {
                    c = target_sample(g, c, ring.peek(0), ring.peek(1), u);
                    used = 2;
                    if (c >= 0) {
                        c = target_sample(g, c, ring.peek(2), ring.peek(3), u);
                        used = 4;
                    }
                }
            }
            cid = __shfl_sync(kFull, c, 0);
            ring.advance(__shfl_sync(kFull, used, 0));
            if (cid < 0) {  // sink: the reference would index row -1; abandon the sample
                dead = true;
                break;
            }
            const int cfield = __ldg(g.field + cid);
            // caller's negative: NegativeSample until same field as the item (HBPR.cpp:110-112)
            int64_t jid[5];
            jid[0] = reject_sample(ring, lane, 2u,
                                   [&](uint32_t w0, uint32_t w1) { return (int64_t)negative_sample(g, w0, w1); },
                                   [&](int64_t x) { return __ldg(g.field + x) == cfield; });
            // rounds 1..4: uniform vertex until same field (proNet.cpp:1477-1482), 1 word per attempt
#pragma unroll
            for (int r = 1; r < 5; ++r)
                jid[r] = reject_sample(ring, lane, 1u,
                                       [&](uint32_t w0, uint32_t) { return (int64_t)index_draw(w0, V32); },
                                       [&](int64_t x) { return __ldg(g.field + x) == cfield; });
            const T alpha = (T)(st.alpha / (double)hop);  // _alpha/w, margin/w (HBPR.cpp:113)
            const T margin = (T)(1.0 / (double)hop);
            const T cdec = A::mul(alpha, (T)0.0025);
            const T cv = A::mul(alpha, (T)0.025);
            T* pv = W + vid * dim;
            T* pi = W + cid * dim;
            bool dup = vid == cid;
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                dup = dup || jid[r] == vid || jid[r] == cid;
#pragma unroll
                for (int s = 0; s < r; ++s) dup = dup || jid[r] == jid[s];
            }
            Row<C> verr;
            verr.zero();
            T up = 0;
            if constexpr (kAtomicRows<C>) {
                // fp32 tables: deltas with red.global.add, the item row's running value tracked in registers
                Row<C> v, ri, rj[5], di;
                v.load_ca(pv, lane, dim);
                ri.load_ca(pi, lane, dim);
#pragma unroll
                for (int r = 0; r < 5; ++r) rj[r].load_ca(W + jid[r] * dim, lane, dim);
                pin(v);
                pin(ri);
#pragma unroll
                for (int r = 0; r < 5; ++r) pin(rj[r]);
                di.zero();
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    Row<C> cvec;
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) cvec.x[e] = A::sub(ri.x[e], rj[r].x[e]);
                    const T f = dot(v, cvec);
                    if (!(f > margin)) {
                        const T gg = A::mul(fast_sigmoid<T>(lut, A::sub((T)0, f)), alpha);
                        up += (T)1;
#pragma unroll
                        for (int e = 0; e < C::EPL; ++e) {
                            verr.x[e] = A::madd(verr.x[e], gg, cvec.x[e]);
                            const T cerr = A::mul(gg, v.x[e]);
                            const T d = A::msub(cerr, cdec, ri.x[e]);
                            di.x[e] = A::add(di.x[e], d);
                            ri.x[e] = A::add(ri.x[e], d);
                            rj[r].x[e] = A::sub(A::mul(-cdec, rj[r].x[e]), cerr);
                        }
                        row_red_add<C>(W + jid[r] * dim, rj[r], lane, dim);
                    }
                }
                if (up > (T)0) {
                    row_red_add<C>(pi, di, lane, dim);
                    const T inv = A::div((T)1, up);
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) v.x[e] = A::msub(A::mul(verr.x[e], inv), cv, v.x[e]);
                    row_red_add<C>(pv, v, lane, dim);
                }
            } else if (!dup) {
                Row<C> v, ri, rj[5];
                v.load_ca(pv, lane, dim);
                ri.load_ca(pi, lane, dim);
#pragma unroll
                for (int r = 0; r < 5; ++r) rj[r].load_ca(W + jid[r] * dim, lane, dim);
                pin(v);
                pin(ri);
#pragma unroll
                for (int r = 0; r < 5; ++r) pin(rj[r]);  // all seven gathers in flight before the first round
                bool ri_dirty = false;
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    Row<C> cvec;
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) cvec.x[e] = A::sub(ri.x[e], rj[r].x[e]);
                    const T f = dot(v, cvec);
                    if (!(f > margin)) {
                        const T gg = A::mul(fast_sigmoid<T>(lut, A::sub((T)0, f)), alpha);
                        up += (T)1;
#pragma unroll
                        for (int e = 0; e < C::EPL; ++e) {
                            verr.x[e] = A::madd(verr.x[e], gg, cvec.x[e]);
                            const T cerr = A::mul(gg, v.x[e]);
                            ri.x[e] = A::msub(ri.x[e], cdec, ri.x[e]);
                            rj[r].x[e] = A::msub(rj[r].x[e], cdec, rj[r].x[e]);
                            ri.x[e] = A::add(ri.x[e], cerr);
                            rj[r].x[e] = A::sub(rj[r].x[e], cerr);
                        }
                        rj[r].store(W + jid[r] * dim, lane, dim);
                        ri_dirty = true;
                    }
                }
                if (ri_dirty) ri.store(pi, lane, dim);
                if (up > (T)0) {
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) {
                        v.x[e] = A::msub(v.x[e], cv, v.x[e]);
                        v.x[e] = A::add(v.x[e], A::div(verr.x[e], up));
                    }
                    v.store(pv, lane, dim);
                }
            } else {
                for (int r = 0; r < 5; ++r)
                    if (ordered_round<C>(pv, pi, W + jid[r] * dim, dim, lane, lut, alpha, true, margin, verr)) up += (T)1;
                if (up > (T)0) {
                    for_owned<C>(lane, dim, [&](int e, int idx) {
                        stv(pv + idx, A::msub(ldv(pv + idx), cv, ldv(pv + idx)));
                        stv(pv + idx, A::add(ldv(pv + idx), A::div(verr.x[e], up)));
                    });
                }
            }
            st.pairs += 5;
        }
        st.count++;
        sched_tick(st, a.sched);
    }
    st.pos = ring.pos;
    if (lane == 0) a.state[w] = st;
}

// HOP-Rec, fp32 Hogwild throughput path (same idea as k_warp_fast / k_hpe_fast). k_hoprec consumes its stream word for word:
// per hop the item's neighbour chain, then five rejection loops one after the other (each has to know how many words the
// previous one used), then the rows -- about 90 dependent memory round trips per sample. Here every sample owns a FIXED
// slice of its warp's stream,
//   [0, 32)    user: up to 16 alias attempts (p, idx) until field 0            (HBPR.cpp:97-99)
//   [32, 52)   the walk: hop 1 one TargetSample, hops 2..5 two each            (HBPR.cpp:100, 105-109)
//   then per hop 192 words: 32 alias attempts (idx, p) for the caller's negative (field of the item, HBPR.cpp:110-112)
//                           and 4 x 32 one-word attempts for the uniform negatives of rounds 1..4 (proNet.cpp:1477-1482)
// (a rejection loop that finds nothing in its 32 attempts -- probability (5/6)^32 on a 5 : 1 user : item mix -- continues on
// an overflow stream), which lets 8 lanes draw the user and the whole item chain of 8 consecutive samples at once and makes
// the 25 rejection loops of a sample independent of each other: the five of a hop are resolved together, their Philox
// blocks computed once per warp and handed round by shuffles.
constexpr int kHopGroup = 8;
constexpr int kHopMaxSteps = 5;
constexpr uint32_t kHopUserWords = 32, kHopWalkOff = 32, kHopNegOff = 52, kHopHopWords = 192;
__host__ __device__ constexpr uint32_t hop_slice_words(int steps) { return kHopNegOff + kHopHopWords * (uint32_t)steps; }
constexpr uint64_t kHopOverflowStream = 1ull << 40;

template <class C>
__global__ void __launch_bounds__(kBlockThreads, walk_min_blocks<C>()) k_hoprec_fast(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    using A = Ar<T>;
    const T* lut = stage_lut<T>(a.lut, reinterpret_cast<T*>(smem_raw));
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (w >= a.n_warps) return;
    WarpState st = a.state[w];
    const GraphDev& g = a.g;
    const int dim = a.dim, steps = a.steps;
    T* W = a.Wv;
    const uint32_t V32 = (uint32_t)g.V;
    const uint64_t stream = a.stream_base + (uint64_t)w;
    const uint64_t slice_blocks = (uint64_t)(hop_slice_words(steps) >> 2);
    const uint64_t s_base = (st.pos + hop_slice_words(steps) - 1) / hop_slice_words(steps);
    for (uint64_t it0 = 0; it0 < a.jobs; it0 += kHopGroup) {
        const int ng = (int)min((uint64_t)kHopGroup, a.jobs - it0);
        // ---- 8 lanes: user + item chain of 8 samples ----
        int m_user = -1, m_item[kHopMaxSteps], m_field[kHopMaxSteps];
#pragma unroll
        for (int h = 0; h < kHopMaxSteps; ++h) m_item[h] = -1, m_field[h] = 0;
        if (lane < ng) {
            const uint64_t sidx = s_base + it0 + (uint64_t)lane;
            const uint64_t blk = slice_blocks * sidx;
            for (uint32_t att = 0; m_user < 0; ++att) {
                U4 r;
                if (att < kHopUserWords / 2) r = philox_block(a.seed, stream, blk + (att >> 1));
                else r = philox_block(a.seed, stream + kHopOverflowStream, sidx * 4096u + (att >> 1));
                const int c = (int)source_sample(g, (att & 1) ? r.z : r.x, (att & 1) ? r.w : r.y);
                if (__ldg(g.field + c) == 0) m_user = c;
            }
            int64_t c = m_user;
            int used;
            U4 r = philox_block(a.seed, stream, blk + (kHopWalkOff >> 2));
            c = target_sample(g, c, r.x, r.y, used);
            m_item[0] = (int)c;
            // hops 2..5: words 2 + 4 (h - 1) .. + 3 of the walk region, i.e. (z, w) of one block and (x, y) of the next
#pragma unroll
            for (int h = 1; h < kHopMaxSteps; ++h) {
                if (h < steps && c >= 0) {
                    const U4 r2 = philox_block(a.seed, stream, blk + (kHopWalkOff >> 2) + (uint32_t)h);
                    c = target_sample(g, c, r.z, r.w, used);
                    if (c >= 0) c = target_sample(g, c, r2.x, r2.y, used);
                    r = r2;
                    m_item[h] = (int)c;
                }
            }
#pragma unroll
            for (int h = 0; h < kHopMaxSteps; ++h)
                if (m_item[h] >= 0) m_field[h] = __ldg(g.field + m_item[h]);
        }
        for (int k = 0; k < ng; ++k) {
            const int vid = __shfl_sync(kFull, m_user, k);
            int idist = -1, fdist = 0;  // lane h: the item of hop h + 1 and its field
#pragma unroll
            for (int h = 0; h < kHopMaxSteps; ++h) {
                const int c = __shfl_sync(kFull, m_item[h], k);
                const int f = __shfl_sync(kFull, m_field[h], k);
                if (lane == h) idist = c, fdist = f;
            }
            const uint64_t sidx = s_base + it0 + (uint64_t)k;
            const uint64_t blk = slice_blocks * sidx;
            T* pv = W + (size_t)vid * dim;
            // the five negatives of hop `hop` (item field cfield); jid[r] is warp-uniform
            auto resolve = [&](int hop, int cfield, int (&jid)[5]) {
                // the hop's 192 words = 48 Philox blocks: lane L computes blocks L and (L < 16) 32 + L
                const uint64_t hblk = blk + (kHopNegOff >> 2) + (uint64_t)(hop - 1) * (kHopHopWords >> 2);
                const U4 ra = philox_block(a.seed, stream, hblk + (uint64_t)lane);
                U4 rb = ra;
                if (lane < 16) rb = philox_block(a.seed, stream, hblk + 32u + (uint64_t)lane);
                // region 0 (words 0..63, blocks 0..15): attempt t = words 2t, 2t+1 = block t/2, (x,y) or (z,w)
                // regions 1..4 (words 64 + 32 (r-1) + t, blocks 16 + 8 (r-1) + t/4): one word per attempt
                int cand[5];
                {
                    const int src = lane >> 1;
                    const uint32_t x = __shfl_sync(kFull, ra.x, src), y = __shfl_sync(kFull, ra.y, src);
                    const uint32_t z = __shfl_sync(kFull, ra.z, src), q = __shfl_sync(kFull, ra.w, src);
                    cand[0] = (int)negative_sample(g, (lane & 1) ? z : x, (lane & 1) ? q : y);
                }
#pragma unroll
                for (int r = 1; r < 5; ++r) {
                    const int b = 16 + 8 * (r - 1) + (lane >> 2);  // block index inside the hop: < 32 lives in ra, else in rb
                    const int src = b & 31;
                    const uint32_t xa = __shfl_sync(kFull, ra.x, src), ya = __shfl_sync(kFull, ra.y, src);
                    const uint32_t za = __shfl_sync(kFull, ra.z, src), qa = __shfl_sync(kFull, ra.w, src);
                    const uint32_t xb = __shfl_sync(kFull, rb.x, src), yb = __shfl_sync(kFull, rb.y, src);
                    const uint32_t zb = __shfl_sync(kFull, rb.z, src), qb = __shfl_sync(kFull, rb.w, src);
                    const bool hi = b >= 32;
                    const int e = lane & 3;
                    const uint32_t wd = e == 0 ? (hi ? xb : xa) : e == 1 ? (hi ? yb : ya) : e == 2 ? (hi ? zb : za) : (hi ? qb : qa);
                    cand[r] = (int)index_draw(wd, V32);
                }
                int fld[5];
#pragma unroll
                for (int r = 0; r < 5; ++r) fld[r] = __ldg(g.field + cand[r]);  // five independent look-ups in flight
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    unsigned acc = __ballot_sync(kFull, fld[r] == cfield);
                    int cnd = cand[r];
                    for (uint32_t round = 1; !acc; ++round) {  // (rare) nothing in 32 attempts: overflow stream
                        const U4 ro = philox_block(a.seed, stream + kHopOverflowStream,
                                                   (sidx * 64u + (uint64_t)(hop * 8 + r)) * 4096u + 2048u + round * 32u + (uint32_t)lane);
                        cnd = r == 0 ? (int)negative_sample(g, ro.x, ro.y) : (int)index_draw(ro.x, V32);
                        acc = __ballot_sync(kFull, __ldg(g.field + cnd) == cfield);
                    }
                    jid[r] = __shfl_sync(kFull, cnd, __ffs(acc) - 1);
                }
            };
            int jid[5], jnext[5];
            int cid = __shfl_sync(kFull, idist, 0);
            if (cid >= 0) resolve(1, __shfl_sync(kFull, fdist, 0), jid);
            for (int hop = 1; hop <= steps; ++hop) {
                if (cid < 0) break;  // sink: the reference would index row -1; abandon the sample
                const T alpha = (T)(st.alpha / (double)hop);  // _alpha/w, margin/w (HBPR.cpp:113)
                const T margin = (T)(1.0 / (double)hop);
                const T cdec = A::mul(alpha, (T)0.0025);
                const T cv = A::mul(alpha, (T)0.025);
                T* pi = W + (size_t)cid * dim;
                Row<C> v, ri, rj[5], di, verr;
                v.load_ca(pv, lane, dim);
                ri.load_ca(pi, lane, dim);
#pragma unroll
                for (int r = 0; r < 5; ++r) rj[r].load_ca(W + (size_t)jid[r] * dim, lane, dim);
                pin(v);
                pin(ri);
#pragma unroll
                for (int r = 0; r < 5; ++r) pin(rj[r]);
                // while the seven rows are in flight: the ids of the next hop
                const int cnext = hop < steps ? __shfl_sync(kFull, idist, hop) : -1;
                if (cnext >= 0) resolve(hop + 1, __shfl_sync(kFull, fdist, hop), jnext);
                di.zero();
                verr.zero();
                T up = 0;
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    Row<C> cvec;
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) cvec.x[e] = A::sub(ri.x[e], rj[r].x[e]);
                    const T f = dot(v, cvec);
                    if (!(f > margin)) {
                        const T gg = A::mul(fast_sigmoid<T>(lut, A::sub((T)0, f)), alpha);
                        up += (T)1;
#pragma unroll
                        for (int e = 0; e < C::EPL; ++e) {
                            verr.x[e] = A::madd(verr.x[e], gg, cvec.x[e]);
                            const T cerr = A::mul(gg, v.x[e]);
                            const T d = A::msub(cerr, cdec, ri.x[e]);
                            di.x[e] = A::add(di.x[e], d);
                            ri.x[e] = A::add(ri.x[e], d);
                            rj[r].x[e] = A::sub(A::mul(-cdec, rj[r].x[e]), cerr);
                        }
                        row_red_add<C>(W + (size_t)jid[r] * dim, rj[r], lane, dim);
                    }
                }
                if (up > (T)0) {
                    row_red_add<C>(pi, di, lane, dim);
                    const T inv = A::div((T)1, up);
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) v.x[e] = A::msub(A::mul(verr.x[e], inv), cv, v.x[e]);
                    row_red_add<C>(pv, v, lane, dim);
                }
                st.pairs += 5;
                cid = cnext;
#pragma unroll
                for (int r = 0; r < 5; ++r) jid[r] = jnext[r];
            }
            st.count++;
            sched_tick(st, a.sched);
        }
    }
    st.pos = (s_base + a.jobs) * hop_slice_words(steps);
    if (lane == 0) a.state[w] = st;
}

// ---------------------------------------------------------------------------------------------------------------
// CPR and TPR (Go tree only): pairwise ranking where one side of the score is AGGREGATED on the fly from the rows of a
// vertex' neighbours, in a second graph as well as in the first. One warp per sample, lane-parallel over the dimension:
// every per-dimension sum runs in the reference's order (adjacency order), so the fp64 deterministic mode reproduces the
// restatements bit for bit; the fp32 Hogwild mode pushes row deltas with red.global.add like every other trainer.
// ---------------------------------------------------------------------------------------------------------------
// acc += sum of tab[col[e]] for e in [beg, end), four gathers in flight; returns the number of rows added
template <class C>
__device__ __forceinline__ int add_neighbour_rows(Row<C>& acc, const typename C::T* tab, const int32_t* __restrict__ col, int64_t beg,
                                                  int64_t end, int dim, int lane) {
    using A = Ar<typename C::T>;
    for (int64_t e0 = beg; e0 < end; e0 += 32) {
        const int nb = (int)min((int64_t)32, end - e0);
        const int my = lane < nb ? __ldg(col + e0 + lane) : 0;
        for (int r0 = 0; r0 < nb; r0 += 4) {
            Row<C> rr[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int id = __shfl_sync(kFull, my, (r0 + r) & 31);
                if (r0 + r < nb) rr[r].load(tab + (size_t)id * dim, lane, dim);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (r0 + r < nb) {
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) acc.x[e] = A::add(acc.x[e], rr[r].x[e]);
                }
        }
    }
    return (int)(end - beg);
}

// CPR.Train (internal/models/cpr/cpr.go:175-282) with transformUser (:127-172). Wv = user rows, Wc = target-domain item rows,
// aux_tab = source-domain item rows. lambda = user_reg. A sample whose user has no target-domain neighbour is skipped
// without counting (:214-216).
// FAST (fp32 Hogwild): the (user, positive, negative) triples of 8 consecutive samples are drawn by 8 lanes at once from fixed
// 8-word slices of the warp's stream (source 2 words, target 1, negative 2) instead of one after the other by lane 0.
template <class C, bool FAST = false>
__global__ void __launch_bounds__(kBlockThreads) k_cpr(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    using A = Ar<T>;
    uint32_t* rings = reinterpret_cast<uint32_t*>(smem_raw);
    T* lut_s = reinterpret_cast<T*>(smem_raw + kWarpsPerBlock * 256 * sizeof(uint32_t));
    const T* lut = stage_lut<T>(a.lut, lut_s);
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= a.n_warps) return;
    WarpState st = a.state[w];
    DrawRing ring;
    ring.init(rings + wib * 256, a.seed, a.stream_base + (uint64_t)w, st.pos, lane);
    const GraphDev& g = a.g;
    const int dim = a.dim;
    const uint64_t fast_base = (st.pos + 7) / 8;  // FAST: samples this stream has already served (8-word slices)
    int64_t gu = -1, gp = -1, gn = -1;
    for (uint64_t it = 0; it < a.jobs; ++it) {
        int64_t u = -1, p = -1, n = -1;
        if constexpr (FAST) {
            if ((it & 7) == 0) {  // lanes 0..7 draw the triples of samples it .. it + 7
                gu = gp = gn = -1;
                if (lane < 8 && it + (uint64_t)lane < a.jobs) {
                    const uint64_t blk = 2 * (fast_base + it + (uint64_t)lane);
                    const U4 r0 = philox_block(a.seed, a.stream_base + (uint64_t)w, blk);
                    const U4 r1 = philox_block(a.seed, a.stream_base + (uint64_t)w, blk + 1);
                    int used;
                    gu = (int64_t)source_sample(g, r0.x, r0.y);
                    gp = target_sample(g, gu, r0.z, r0.w, used);
                    gn = (int64_t)negative_sample(g, r1.x, r1.y);
                }
            }
            u = __shfl_sync(kFull, gu, (int)(it & 7));
            p = __shfl_sync(kFull, gp, (int)(it & 7));
            n = __shfl_sync(kFull, gn, (int)(it & 7));
            if (p < 0) continue;
        } else {
            ring.ensure();
            int used = 0;
            if (lane == 0) {
                u = (int64_t)source_sample(g, ring.peek(0), ring.peek(1));
                p = target_sample(g, u, ring.peek(2), ring.peek(3), used);
                if (p >= 0) n = (int64_t)negative_sample(g, ring.peek(2u + (uint32_t)used), ring.peek(3u + (uint32_t)used));
            }
            u = __shfl_sync(kFull, u, 0);
            p = __shfl_sync(kFull, p, 0);
            n = __shfl_sync(kFull, n, 0);
            used = __shfl_sync(kFull, used, 0);
            if (p < 0) {
                ring.advance(2u + (uint32_t)used);
                continue;
            }
            ring.advance(4u + (uint32_t)used);
        }
        const T alpha = (T)st.alpha;
        T* pu = a.Wv + (size_t)u * dim;
        T* pp = a.Wc + (size_t)p * dim;
        T* pn = a.Wc + (size_t)n * dim;
        Row<C> ru, rp, rn, uv;
        ru.load(pu, lane, dim);
        rp.load(pp, lane, dim);
        rn.load(pn, lane, dim);
        // transformUser: (user + target-domain neighbours + source-domain neighbours) / count, summed from 0.0 in that order
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) uv.x[e] = A::add((T)0, ru.x[e]);
        int cnt = 1;
        cnt += add_neighbour_rows<C>(uv, a.Wc, g.col, __ldg(g.row_off + u), __ldg(g.row_off + u + 1), dim, lane);
        if (u < a.aux_V) cnt += add_neighbour_rows<C>(uv, a.aux_tab, a.aux_col, __ldg(a.aux_off + u), __ldg(a.aux_off + u + 1), dim, lane);
        const T fc = (T)cnt;
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) uv.x[e] = A::div(uv.x[e], fc);
        Row<C> two[2] = {rp, rn};
        T sc[2];
        dots<C, 2>(uv, two, 2, sc);
        const T diff = A::sub(sc[0], sc[1]);
        if (diff < a.margin) {
            const T gc = A::mul(alpha, fast_sigmoid<T>(lut, A::sub((T)0, A::sub(diff, a.margin))));
            const T cu = A::mul(alpha, a.lambda), ci = A::mul(alpha, a.item_reg);
            if constexpr (kAtomicRows<C>) {
#pragma unroll
                for (int e = 0; e < C::EPL; ++e) {
                    const T pg = A::mul(gc, uv.x[e]);
                    const T ug = A::mul(gc, A::sub(rp.x[e], rn.x[e]));
                    ru.x[e] = A::msub(ug, cu, ru.x[e]);           // -(alpha*user_reg)*w + ug
                    rp.x[e] = A::msub(pg, ci, rp.x[e]);           // -(alpha*item_reg)*w + pg
                    rn.x[e] = A::sub(A::mul(-ci, rn.x[e]), pg);   // -(alpha*item_reg)*w - pg
                }
                row_red_add<C>(pu, ru, lane, dim);
                row_red_add<C>(pp, rp, lane, dim);
                row_red_add<C>(pn, rn, lane, dim);
            } else {
                const bool same = p == n;
#pragma unroll
                for (int e = 0; e < C::EPL; ++e) {
                    const T pg = A::mul(gc, uv.x[e]);
                    const T ng = A::mul(-gc, uv.x[e]);
                    const T ug = A::mul(gc, A::sub(rp.x[e], rn.x[e]));
                    ru.x[e] = A::add(A::msub(ru.x[e], cu, ru.x[e]), ug);
                    rp.x[e] = A::add(A::msub(rp.x[e], ci, rp.x[e]), pg);
                    const T nc = same ? rp.x[e] : rn.x[e];  // pos == neg: the second pair of statements sees the first (:251-258)
                    rn.x[e] = A::add(A::msub(nc, ci, nc), ng);
                }
                ru.store(pu, lane, dim);
                if (!same) rp.store(pp, lane, dim);
                rn.store(pn, lane, dim);
            }
            st.pairs++;
        }
        st.count++;
        sched_tick(st, a.sched);
    }
    st.pos = FAST ? (fast_base + a.jobs) * 8 : ring.pos;
    if (lane == 0) a.state[w] = st;
}

// getTextEnrichedItemEmbedding (tpr.go:101-121): (1 - tw) * item + sum over the item's words of (tw / #words) * word, or the
// item row itself when the item has no words. `ri` returns the item's base row, `nw` its number of words.
template <class C>
__device__ __forceinline__ void tpr_enriched(const TrainArgs<typename C::T>& a, int64_t item, Row<C>& ri, Row<C>& out, int64_t& beg,
                                             int& nw, int lane) {
    using T = typename C::T;
    using A = Ar<T>;
    const int dim = a.dim;
    ri.load(a.Wc + (size_t)item * dim, lane, dim);
    beg = 0;
    nw = 0;
    if (item < a.aux_V) {
        beg = __ldg(a.aux_off + item);
        nw = (int)(__ldg(a.aux_off + item + 1) - beg);
    }
    if (nw <= 0) {
        out = ri;
        return;
    }
    const T base = A::sub((T)1, a.text_w);
    const T wq = A::div(a.text_w, (T)nw);
#pragma unroll
    for (int e = 0; e < C::EPL; ++e) out.x[e] = A::mul(base, ri.x[e]);
    for (int e0 = 0; e0 < nw; e0 += 32) {
        const int nb = min(32, nw - e0);
        const int my = lane < nb ? __ldg(a.aux_col + beg + e0 + lane) : 0;
        for (int r0 = 0; r0 < nb; r0 += 4) {
            Row<C> rr[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int id = __shfl_sync(kFull, my, (r0 + r) & 31);
                if (r0 + r < nb) rr[r].load(a.aux_tab + (size_t)id * dim, lane, dim);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (r0 + r < nb) {
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) out.x[e] = A::madd(out.x[e], wq, rr[r].x[e]);
                }
        }
    }
}

// word rows of one item: w += (tw / #words) * grad - (lambda * alpha) * w, one word after the other (tpr.go:216-232)
template <class C>
__device__ __forceinline__ void tpr_push_words(const TrainArgs<typename C::T>& a, int64_t beg, int nw, const Row<C>& grad,
                                               typename C::T la, int lane) {
    using T = typename C::T;
    using A = Ar<T>;
    if (nw <= 0) return;
    const int dim = a.dim;
    const T ww = A::div(a.text_w, (T)nw);
    for (int e0 = 0; e0 < nw; e0 += 32) {
        const int nb = min(32, nw - e0);
        const int my = lane < nb ? __ldg(a.aux_col + beg + e0 + lane) : 0;
        for (int r = 0; r < nb; ++r) {
            T* pw = a.aux_tab + (size_t)__shfl_sync(kFull, my, r) * dim;
            Row<C> rw;
            rw.load(pw, lane, dim);
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) {
                const T d = A::msub(A::mul(ww, grad.x[e]), la, rw.x[e]);
                rw.x[e] = kAtomicRows<C> ? d : A::add(rw.x[e], d);
            }
            if constexpr (kAtomicRows<C>) row_red_add<C>(pw, rw, lane, dim);
            else rw.store(pw, lane, dim);  // (a repeated word is re-read by the lanes that wrote it: program order)
        }
    }
}

// TPR.Train (internal/models/tpr/tpr.go:124-262). Wv = user rows, Wc = item rows (both over the user-item graph's vids),
// aux_tab = word rows over the item-word graph's vids.
template <class C, bool FAST = false>
__global__ void __launch_bounds__(kBlockThreads) k_tpr(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    using A = Ar<T>;
    uint32_t* rings = reinterpret_cast<uint32_t*>(smem_raw);
    T* lut_s = reinterpret_cast<T*>(smem_raw + kWarpsPerBlock * 256 * sizeof(uint32_t));
    const T* lut = stage_lut<T>(a.lut, lut_s);
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= a.n_warps) return;
    WarpState st = a.state[w];
    DrawRing ring;
    ring.init(rings + wib * 256, a.seed, a.stream_base + (uint64_t)w, st.pos, lane);
    const GraphDev& g = a.g;
    const int dim = a.dim;
    const uint64_t fast_base = (st.pos + 7) / 8;  // FAST: samples this stream has already served (8-word slices)
    int64_t gu = -1, gp = -1, gn = -1;
    for (uint64_t it = 0; it < a.jobs; ++it) {
        int64_t u = -1, p = -1, n = -1;
        if constexpr (FAST) {
            if ((it & 7) == 0) {  // lanes 0..7 draw the triples of samples it .. it + 7
                gu = gp = gn = -1;
                if (lane < 8 && it + (uint64_t)lane < a.jobs) {
                    const uint64_t blk = 2 * (fast_base + it + (uint64_t)lane);
                    const U4 r0 = philox_block(a.seed, a.stream_base + (uint64_t)w, blk);
                    const U4 r1 = philox_block(a.seed, a.stream_base + (uint64_t)w, blk + 1);
                    int used;
                    gu = (int64_t)source_sample(g, r0.x, r0.y);
                    gp = target_sample(g, gu, r0.z, r0.w, used);
                    gn = (int64_t)negative_sample(g, r1.x, r1.y);
                }
            }
            u = __shfl_sync(kFull, gu, (int)(it & 7));
            p = __shfl_sync(kFull, gp, (int)(it & 7));
            n = __shfl_sync(kFull, gn, (int)(it & 7));
            if (p < 0) continue;
        } else {
            ring.ensure();
            int used = 0;
            if (lane == 0) {
                u = (int64_t)source_sample(g, ring.peek(0), ring.peek(1));
                p = target_sample(g, u, ring.peek(2), ring.peek(3), used);
                if (p >= 0) n = (int64_t)negative_sample(g, ring.peek(2u + (uint32_t)used), ring.peek(3u + (uint32_t)used));
            }
            u = __shfl_sync(kFull, u, 0);
            p = __shfl_sync(kFull, p, 0);
            n = __shfl_sync(kFull, n, 0);
            used = __shfl_sync(kFull, used, 0);
            if (p < 0) {
                ring.advance(2u + (uint32_t)used);
                continue;
            }
            ring.advance(4u + (uint32_t)used);
        }
        const T alpha = (T)st.alpha;
        const T la = A::mul(a.lambda, alpha);
        T* pu = a.Wv + (size_t)u * dim;
        T* pp = a.Wc + (size_t)p * dim;
        T* pn = a.Wc + (size_t)n * dim;
        Row<C> ru, rp, rn, two[2];
        int64_t pbeg, nbeg;
        int pnw, nnw;
        ru.load(pu, lane, dim);
        tpr_enriched<C>(a, p, rp, two[0], pbeg, pnw, lane);
        tpr_enriched<C>(a, n, rn, two[1], nbeg, nnw, lane);
        T sc[2];
        dots<C, 2>(ru, two, 2, sc);
        const T gc = A::mul(alpha, fast_sigmoid<T>(lut, A::sub(sc[1], sc[0])));
        const T base = A::sub((T)1, a.text_w);
        Row<C> pg, ng;
        const bool same = p == n;
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const T ug = A::mul(gc, A::sub(two[0].x[e], two[1].x[e]));
            pg.x[e] = A::mul(gc, ru.x[e]);
            ng.x[e] = A::mul(-gc, ru.x[e]);
            const T du = A::msub(ug, la, ru.x[e]);
            const T dp = A::msub(A::mul(base, pg.x[e]), la, rp.x[e]);
            if constexpr (kAtomicRows<C>) {
                ru.x[e] = du;
                rp.x[e] = dp;
                rn.x[e] = A::msub(A::mul(base, ng.x[e]), la, rn.x[e]);
            } else {
                ru.x[e] = A::add(ru.x[e], du);
                rp.x[e] = A::add(rp.x[e], dp);
                const T nc = same ? rp.x[e] : rn.x[e];  // pos == neg: the negative's statement sees the positive's (:211-212)
                rn.x[e] = A::add(nc, A::msub(A::mul(base, ng.x[e]), la, nc));
            }
        }
        if constexpr (kAtomicRows<C>) {
            row_red_add<C>(pu, ru, lane, dim);
            row_red_add<C>(pp, rp, lane, dim);
            row_red_add<C>(pn, rn, lane, dim);
        } else {
            ru.store(pu, lane, dim);
            if (!same) rp.store(pp, lane, dim);
            rn.store(pn, lane, dim);
        }
        tpr_push_words<C>(a, pbeg, pnw, pg, la, lane);
        tpr_push_words<C>(a, nbeg, nnw, ng, la, lane);
        st.pairs++;
        st.count++;
        sched_tick(st, a.sched);
    }
    st.pos = FAST ? (fast_base + a.jobs) * 8 : ring.pos;
    if (lane == 0) a.state[w] = st;
}

}  // namespace smore
