// Device-side graph construction:
//
//   * alias_build_device   Walker/Vose alias tables built IN PARALLEL on the GPU. The reference's AliasMethod
//                          (src/proNet.cpp:544-620) is a sequential two-stack sweep -- 4 G pow() + Vose steps on one host core
//                          at the north-star size. The parallel construction is the "split" formulation of the same sweep
//                          (Huebschle-Schneider & Sanders, "Parallel Weighted Random Sampling"): with the light items l_0..
//                          (scaled weight p < 1) and the heavy items h_0.. in index order, SL(i) = sum_{k<i} (1 - p[l_k]) and
//                          SH(j) = sum_{k<=j} (p[h_k] - 1), the sweep gives light l_i the alias h_j with j = min{j : SH(j) >=
//                          SL(i)}, and heavy h_j ends as a bucket of probability 1 + SH(j) - SL(i*) with alias h_{j+1}, i* =
//                          min{i : SL(i) > SH(j)} -- two prefix sums and two binary searches per item, no sequential step.
//                          The table differs from the reference's (other pairing order) but encodes the same distribution
//                          exactly (tests/test_gpu_device_graph.py); it is used where no reference table exists to compare
//                          with: the per-block edge tables, shard-local negative tables and sub-part vertex tables of the
//                          row-sharded store.
//   * smore_graph_create_synthetic_rotating
//                          the bench graph of BASELINE configs[4] shape (V vertices, E undirected weighted edge lines,
//                          power-law endpoints) generated straight into the per-rank block tables of the rotating-shard mode:
//                          the edge list is a pure function of (seed, line index), every rank scans it and keeps the entries
//                          whose target it owns; nothing goes through the host. The reference cannot load such a graph at all
//                          (30 M-vertex hash limit src/proNet.h:34, int-narrowed random_gen src/random.cpp:5).
#include "host_common.h"

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;                         // per thread
constexpr int kScanChunk = kScanThreads * kScanItems;  // per block

struct Tri {
    long long nl;  // light items
    double dl;     // deficits of the light items: 1 - p
    double dh;     // excesses of the heavy items: p - 1
};
__device__ __forceinline__ Tri tri_add(const Tri& a, const Tri& b) { return Tri{a.nl + b.nl, a.dl + b.dl, a.dh + b.dh}; }
__device__ __forceinline__ Tri tri_of(double p) {
    return p < 1.0 ? Tri{1, 1.0 - p, 0.0} : Tri{0, 0.0, p - 1.0};
}
__device__ __forceinline__ Tri tri_shfl_up(const Tri& a, int d) {
    return Tri{__shfl_up_sync(kFull, a.nl, d), __shfl_up_sync(kFull, a.dl, d), __shfl_up_sync(kFull, a.dh, d)};
}

// Exclusive scan of one Tri per thread over the block; returns the thread's prefix, *total = block total.
__device__ Tri block_exscan(Tri mine, Tri* total) {
    __shared__ Tri wsum[kScanThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Tri inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Tri o = tri_shfl_up(inc, d);
        if (lane >= d) inc = tri_add(o, inc);
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    Tri base{0, 0.0, 0.0};
    Tri tot{0, 0.0, 0.0};
    for (int w = 0; w < kScanThreads / 32; ++w) {
        if (w == wid) base = tot;
        tot = tri_add(tot, wsum[w]);
    }
    __syncthreads();
    *total = tot;
    Tri prev = tri_shfl_up(inc, 1);
    if (lane == 0) prev = Tri{0, 0.0, 0.0};
    return tri_add(base, prev);
}

__global__ void k_sum(const double* __restrict__ w, long long n, double* out) {
    double s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += w[i];
    s = warp_sum(s);
    __shared__ double ws[32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) atomicAdd(out, t);
    }
}

// phase 1: per-chunk totals of (light count, deficit sum, excess sum); p = w * scale
__global__ void __launch_bounds__(kScanThreads) k_alias_reduce(const double* __restrict__ w, long long n, double scale, Tri* bsum) {
    const long long base = (long long)blockIdx.x * kScanChunk + (long long)threadIdx.x * kScanItems;
    Tri t{0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) t = tri_add(t, tri_of(w[base + k] * scale));
    Tri tot;
    block_exscan(t, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// phase 2: exclusive scan of the chunk totals (one block), grand total appended at bsum[nb]
__global__ void __launch_bounds__(kScanThreads) k_alias_scan_bsums(Tri* bsum, long long nb) {
    __shared__ Tri carry;
    if (threadIdx.x == 0) carry = Tri{0, 0.0, 0.0};
    __syncthreads();
    for (long long b0 = 0; b0 < nb; b0 += kScanThreads) {
        const long long i = b0 + threadIdx.x;
        Tri mine = i < nb ? bsum[i] : Tri{0, 0.0, 0.0};
        Tri tot;
        Tri pre = block_exscan(mine, &tot);
        const Tri c = carry;
        if (i < nb) bsum[i] = tri_add(c, pre);
        __syncthreads();
        if (threadIdx.x == 0) carry = tri_add(c, tot);
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[nb] = carry;
}

// phase 3: scatter -- light item #i: L[i] = index, SL[i] = deficits before it; heavy item #j: H[j] = index,
// SH[j] = excesses up to and including it
__global__ void __launch_bounds__(kScanThreads) k_alias_scatter(const double* __restrict__ w, long long n, double scale,
                                                                const Tri* __restrict__ bsum, uint32_t* L, double* SL,
                                                                uint32_t* H, double* SH) {
    const long long base = (long long)blockIdx.x * kScanChunk + (long long)threadIdx.x * kScanItems;
    double p[kScanItems];
    Tri t{0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        p[k] = base + k < n ? w[base + k] * scale : 1.0;
        if (base + k < n) t = tri_add(t, tri_of(p[k]));
    }
    Tri tot;
    Tri pre = tri_add(bsum[blockIdx.x], block_exscan(t, &tot));
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k >= n) break;
        const long long idx = base + k;
        if (p[k] < 1.0) {
            L[pre.nl] = (uint32_t)idx;
            SL[pre.nl] = pre.dl;
            pre.nl += 1;
            pre.dl += 1.0 - p[k];
        } else {
            const long long j = idx - pre.nl;  // heavy items before idx = idx - light items before idx
            pre.dh += p[k] - 1.0;
            H[j] = (uint32_t)idx;
            SH[j] = pre.dh;
        }
    }
}

__device__ __forceinline__ uint2 pack_entry(double prob, uint32_t alias, uint32_t self) {
    const double scaled = ceil(prob * 4294967296.0);
    if (!(scaled < 4294967296.0)) return make_uint2(0xFFFFFFFFu, self);
    return make_uint2(scaled <= 0.0 ? 0u : (uint32_t)scaled, alias);
}

// light l_i -> alias h_j, j = first heavy with SH[j] >= SL[i]
__global__ void k_alias_lights(const double* __restrict__ w, double scale, const uint32_t* __restrict__ L,
                               const double* __restrict__ SL, long long a, const uint32_t* __restrict__ H,
                               const double* __restrict__ SH, long long b, uint2* out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t idx = L[i];
        const double need = SL[i];
        long long lo = 0, hi = b;  // first j in [0, b) with SH[j] >= need, b if none
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (SH[mid] >= need) hi = mid;
            else lo = mid + 1;
        }
        // (no heavy left: only by rounding at the very end of the sweep -- the bucket keeps itself)
        out[idx] = lo < b ? pack_entry(w[idx] * scale, H[lo], idx) : make_uint2(0xFFFFFFFFu, idx);
    }
}

// heavy h_j -> probability 1 + SH[j] - SL[i*], alias h_{j+1}; i* = first light count with SL(i*) > SH[j] (SL(a) = total deficit)
__global__ void k_alias_heavies(const uint32_t* __restrict__ H, const double* __restrict__ SH, long long b,
                                const double* __restrict__ SL, long long a, double total_deficit, uint2* out) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < b; j += (long long)gridDim.x * blockDim.x) {
        const uint32_t idx = H[j];
        const double have = SH[j];
        long long lo = 0, hi = a + 1;  // first i in [0, a] with SLx(i) > have, a + 1 if none; SLx(a) = total_deficit
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            const double v = mid < a ? SL[mid] : total_deficit;
            if (v > have) hi = mid;
            else lo = mid + 1;
        }
        if (lo > a || j + 1 >= b) {
            out[idx] = make_uint2(0xFFFFFFFFu, idx);  // never exhausted (the last heavy, up to rounding): probability 1
        } else {
            const double v = lo < a ? SL[lo] : total_deficit;
            double prob = 1.0 + have - v;
            prob = prob < 0.0 ? 0.0 : prob;
            out[idx] = pack_entry(prob, H[j + 1], idx);
        }
    }
}

__global__ void k_alias_uniform(long long n, uint2* out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = make_uint2(0xFFFFFFFFu, (uint32_t)i);
}

struct AliasScratchDev {
    DevBuf L, H, SL, SH, bsum, sum;
};

// out[i] = {threshold, alias} for weights w[0, n) (device pointers; w >= 0). All-zero weights -> uniform (alias.go:30-37).
int alias_build_device(const double* w, int64_t n, uint2* out, AliasScratchDev& s, double* total_out = nullptr) {
    if (n <= 0) return SMORE_OK;
    if (n >= (1ll << 32)) return fail(SMORE_E_UNSUPPORTED, "alias table of %lld entries: 32-bit draws", (long long)n);
    if (int rc = s.sum.ensure(sizeof(double))) return rc;
    CU(cudaMemsetAsync(s.sum.p, 0, sizeof(double)));
    const int grid = 148 * 8;
    k_sum<<<grid, 256>>>(w, n, (double*)s.sum.p);
    g_launches++;
    double total = 0;
    CU(cudaMemcpy(&total, s.sum.p, sizeof(double), cudaMemcpyDeviceToHost));
    if (total_out) *total_out = total;
    if (!(total > 0)) {
        k_alias_uniform<<<grid, 256>>>(n, out);
        g_launches++;
        CU(cudaGetLastError());
        return SMORE_OK;
    }
    const double scale = (double)n / total;
    const int64_t nb = (n + kScanChunk - 1) / kScanChunk;
    if (int rc = s.bsum.ensure((size_t)(nb + 1) * sizeof(Tri))) return rc;
    if (int rc = s.L.ensure((size_t)n * 4)) return rc;
    if (int rc = s.H.ensure((size_t)n * 4)) return rc;
    if (int rc = s.SL.ensure((size_t)n * 8)) return rc;
    if (int rc = s.SH.ensure((size_t)n * 8)) return rc;
    k_alias_reduce<<<(unsigned)nb, kScanThreads>>>(w, n, scale, (Tri*)s.bsum.p);
    k_alias_scan_bsums<<<1, kScanThreads>>>((Tri*)s.bsum.p, nb);
    k_alias_scatter<<<(unsigned)nb, kScanThreads>>>(w, n, scale, (const Tri*)s.bsum.p, (uint32_t*)s.L.p, (double*)s.SL.p,
                                                    (uint32_t*)s.H.p, (double*)s.SH.p);
    g_launches += 3;
    Tri tot;
    CU(cudaMemcpy(&tot, (const Tri*)s.bsum.p + nb, sizeof(Tri), cudaMemcpyDeviceToHost));
    const long long a = tot.nl, b = n - tot.nl;
    if (a > 0) k_alias_lights<<<grid, 256>>>(w, scale, (const uint32_t*)s.L.p, (const double*)s.SL.p, a, (const uint32_t*)s.H.p,
                                             (const double*)s.SH.p, b, out);
    if (b > 0) k_alias_heavies<<<grid, 256>>>((const uint32_t*)s.H.p, (const double*)s.SH.p, b, (const double*)s.SL.p, a, tot.dl, out);
    g_launches += 2;
    CU(cudaGetLastError());
    return SMORE_OK;
}

// ---- synthetic power-law graph: a pure function of (seed, line index) --------------------------------------------------
constexpr uint64_t kSynthStream = (1ull << 62) + (1ull << 61);  // away from the sampler / init / shuffle streams

struct Synth {
    long long V, E;  // vertices, undirected edge lines
    uint64_t seed;
    int bits;        // 2^bits >= V: domain of the label permutation
};

// bijection on [0, 2^bits) (4-round Feistel on the two halves), cycle-walked into [0, V): hot vertices are not the low ids
__device__ __forceinline__ uint32_t synth_perm(const Synth& s, uint32_t x) {
    const int hb = s.bits >> 1, lb = s.bits - hb;  // high / low half widths
    const uint32_t hmask = (1u << hb) - 1u, lmask = (1u << lb) - 1u;
    do {
        uint32_t hi = x >> lb, lo = x & lmask;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            // alternate which half is mixed so that both widths are respected
            uint32_t f = (lo + 0x9E3779B9u * (uint32_t)(r + 1) + (uint32_t)s.seed) * 0x85EBCA6Bu;
            f ^= f >> 13;
            f *= 0xC2B2AE35u;
            f ^= f >> 16;
            hi = (hi ^ f) & hmask;
            uint32_t g = (hi + 0x7F4A7C15u * (uint32_t)(r + 1) + (uint32_t)(s.seed >> 32)) * 0xCC9E2D51u;
            g ^= g >> 15;
            g *= 0x1B873593u;
            g ^= g >> 16;
            lo = (lo ^ g) & lmask;
        }
        x = (hi << lb) | lo;
    } while ((long long)x >= s.V);
    return x;
}

// Endpoint with P(rank) ~ rank^(-2/3) (the density of smore_b200/synth.py power_law_edges, gamma = 2.5, in closed form:
// inverse CDF rank = V * u^3), then relabelled by the permutation.
__device__ __forceinline__ uint32_t synth_endpoint(const Synth& s, uint32_t k) {
    const double u = ((double)k + 0.5) * (1.0 / 4294967296.0);
    long long r = (long long)((double)s.V * u * u * u);
    if (r >= s.V) r = s.V - 1;
    return synth_perm(s, (uint32_t)r);
}

// line i -> (u, v, w); returns false for a self loop (dropped, as a real edge list would not carry it)
__device__ __forceinline__ bool synth_line(const Synth& s, long long i, uint32_t& u, uint32_t& v, int& w) {
    const U4 r = philox_block(s.seed, kSynthStream, (uint64_t)i);
    u = synth_endpoint(s, r.x);
    v = synth_endpoint(s, r.y);
    w = 1 + (int)__umulhi(r.z, 5u);
    return u != v;
}

__constant__ double c_pow075[6];  // w^0.75 for the integer weights 1..5

// pass 1: weighted degrees and the per-vertex normaliser sum_e w_e^0.75 (undirected: every line feeds both endpoints)
__global__ void k_synth_degrees(Synth s, double* out_deg, double* nrm) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s.E; i += (long long)gridDim.x * blockDim.x) {
        uint32_t u, v;
        int w;
        if (!synth_line(s, i, u, v, w)) continue;
        atomicAdd(out_deg + u, (double)w);
        atomicAdd(out_deg + v, (double)w);
        atomicAdd(nrm + u, c_pow075[w]);
        atomicAdd(nrm + v, c_pow075[w]);
    }
}

// psrc[v] = out_deg^pw (C++ 0.75 / Go 1), pvn[v] = psrc / nrm (the per-entry factor before w^pw), sum -> *src_sum
__global__ void k_synth_psrc(long long V, const double* __restrict__ out_deg, const double* __restrict__ nrm, int cpp,
                             double* psrc, double* src_sum) {
    double s = 0;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
        const double d = out_deg[v];
        const double p = d > 0 ? (cpp ? pow(d, 0.75) : d) : 0.0;
        psrc[v] = p;
        s += p;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) atomicAdd(src_sum, s);
}

struct ShardGeom {
    int world, shift, rank;
    long long sub_cap;
    __device__ __forceinline__ int sub_of(uint32_t x, uint32_t* row) const {
        const int home = (int)(x & (uint32_t)(world - 1));
        const long long l = (long long)(x >> shift);
        const int half = l >= sub_cap ? 1 : 0;
        *row = (uint32_t)(l - (half ? sub_cap : 0));
        return 2 * home + half;
    }
};

// pass 2: entries per block (entry a -> b belongs to this rank if it owns b; block = sub-part of a)
__global__ void k_synth_count(Synth s, ShardGeom gm, unsigned long long* cnt) {
    __shared__ unsigned int sc[2 * kMaxWorld];
    if (threadIdx.x < 2 * kMaxWorld) sc[threadIdx.x] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s.E; i += (long long)gridDim.x * blockDim.x) {
        uint32_t u, v, row;
        int w;
        if (!synth_line(s, i, u, v, w)) continue;
        if ((int)(v & (uint32_t)(gm.world - 1)) == gm.rank) atomicAdd(&sc[gm.sub_of(u, &row)], 1u);
        if ((int)(u & (uint32_t)(gm.world - 1)) == gm.rank) atomicAdd(&sc[gm.sub_of(v, &row)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 2 * kMaxWorld && sc[threadIdx.x]) atomicAdd(cnt + threadIdx.x, (unsigned long long)sc[threadIdx.x]);
}

// pass 3: fill the blocks: probability of the entry under the unsharded samplers (psrc[a] / src_sum / nrm[a] * w^pw; Go:
// nrm = out_deg and pw = 1), source ROW inside its sub-part, context LOCAL row
__global__ void k_synth_fill(Synth s, ShardGeom gm, const double* __restrict__ psrc, const double* __restrict__ nrm,
                             const double* __restrict__ out_deg, int cpp, double inv_src_sum, unsigned long long* fill,
                             double* pe, int32_t* esrc, int32_t* edst) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s.E; i += (long long)gridDim.x * blockDim.x) {
        uint32_t u, v;
        int w;
        if (!synth_line(s, i, u, v, w)) continue;
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
            const uint32_t a = dir ? v : u, b = dir ? u : v;
            if ((int)(b & (uint32_t)(gm.world - 1)) != gm.rank) continue;
            uint32_t row;
            const int q = gm.sub_of(a, &row);
            const unsigned long long k = atomicAdd(fill + q, 1ull);
            const double wp = cpp ? c_pow075[w] : (double)w;
            const double den = cpp ? nrm[a] : out_deg[a];
            pe[k] = psrc[a] * inv_src_sum / den * wp;
            esrc[k] = (int32_t)row;
            edst[k] = (int32_t)(b >> gm.shift);
        }
    }
}

// weights of the shard-local negative table: (in + out degree) of local row l = 2 * out_deg (undirected), to the power 0.75
__global__ void k_synth_neg_weights(long long nl, int world, int rank, const double* __restrict__ out_deg, double* wout) {
    for (long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x; l < nl; l += (long long)gridDim.x * blockDim.x) {
        const double d = 2.0 * out_deg[l * world + rank];
        wout[l] = d > 0 ? pow(d, 0.75) : 0.0;
    }
}
// weights of sub-part q's vertex table: psrc of its rows
__global__ void k_synth_sub_weights(long long n, int q, int shift, long long sub_cap, const double* __restrict__ psrc, double* wout) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x)
        wout[r] = psrc[((r + ((q & 1) ? sub_cap : 0)) << shift) + (q >> 1)];
}

int upload_lut(smore_graph_s* g) {
    double lut[kSigmoidTable + 1];
    float lutf[kSigmoidTable + 1];
    for (int i = 0; i != kSigmoidTable + 1; i++) {
        double x = i * 2.0 * 8.0 / kSigmoidTable - 8.0;
        lut[i] = 1.0 / (1.0 + std::exp(-x));
        lutf[i] = (float)lut[i];
    }
    if (int rc = dev_alloc_copy(&g->d_lut64, lut, kSigmoidTable + 1)) return rc;
    return dev_alloc_copy(&g->d_lut32, lutf, kSigmoidTable + 1);
}

}  // namespace

extern "C" {

// Test / tooling hook: the parallel alias construction on host weights; thr[i] / alias[i] as stored on the device.
int smore_alias_build_device(const double* weights, int64_t n, uint32_t* thr, uint32_t* alias) {
    if (!weights || n <= 0 || !thr || !alias) return fail(SMORE_E_INVALID, "bad argument");
    if (int rc = ensure_device()) return rc;
    double* d_w = nullptr;
    uint2* d_out = nullptr;
    if (int rc = dev_alloc_copy(&d_w, weights, (size_t)n)) return rc;
    cudaError_t e = cudaMalloc((void**)&d_out, (size_t)n * sizeof(uint2));
    if (e != cudaSuccess) {
        cudaFree(d_w);
        return fail(SMORE_E_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
    }
    AliasScratchDev s;
    int rc = alias_build_device(d_w, n, d_out, s);
    std::vector<uint2> h((size_t)n);
    if (!rc && cudaMemcpy(h.data(), d_out, (size_t)n * sizeof(uint2), cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = fail(SMORE_E_CUDA, "read-back failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_w);
    cudaFree(d_out);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        thr[i] = h[(size_t)i].x;
        alias[i] = h[(size_t)i].y;
    }
    return SMORE_OK;
}

int smore_graph_create_synthetic_rotating(int64_t V, int64_t E_lines, uint64_t seed, int semantics, int rank, int world,
                                          smore_graph_t* out) {
    if (!out) return fail(SMORE_E_INVALID, "null argument");
    if (semantics != SMORE_SEM_CPP && semantics != SMORE_SEM_GO) return fail(SMORE_E_INVALID, "bad semantics");
    if (world < 2 || world > kMaxWorld || (world & (world - 1))) return fail(SMORE_E_INVALID, "world must be 2, 4 or 8");
    if (rank < 0 || rank >= world) return fail(SMORE_E_INVALID, "rank out of range");
    if (V < 4ll * world || V >= (1ll << 31)) return fail(SMORE_E_INVALID, "V=%lld out of range", (long long)V);
    if (E_lines <= 0) return fail(SMORE_E_INVALID, "E_lines must be positive");
    if (int rc = ensure_device()) return rc;
    int shift = 0;
    while ((1 << shift) < world) ++shift;
    Synth s{V, E_lines, seed, 2};
    while ((1ll << s.bits) < V) ++s.bits;
    const int nsub = 2 * world;
    const int64_t shard_cap = (V + world - 1) / world, sub_cap = (shard_cap + 1) / 2;
    const int64_t nl = (V - rank + world - 1) / world;
    ShardGeom gm{world, shift, rank, sub_cap};
    const int cpp = semantics == SMORE_SEM_CPP ? 1 : 0;
    const int grid = 148 * 16;
    double p075[6] = {0, 1.0, std::pow(2.0, 0.75), std::pow(3.0, 0.75), std::pow(4.0, 0.75), std::pow(5.0, 0.75)};
    CU(cudaMemcpyToSymbol(c_pow075, p075, sizeof(p075)));

    smore_graph_s* g = new smore_graph_s();
    auto bail = [&](int rc) { delete g; return rc; };
    g->sem = semantics;
    g->neg_method = SMORE_NEG_DEGREES;
    g->V = V;
    g->E = 2 * E_lines;  // CSR entries of the undirected graph (self loops excepted)
    g->n_lines = E_lines;
    g->synthetic = true;
    DevBuf out_deg, nrm, psrc, ssum, cnt, wtmp;
    if (int rc = out_deg.ensure((size_t)V * 8)) return bail(rc);
    if (int rc = nrm.ensure((size_t)V * 8)) return bail(rc);
    if (int rc = psrc.ensure((size_t)V * 8)) return bail(rc);
    if (int rc = ssum.ensure(8)) return bail(rc);
    if (int rc = cnt.ensure(2 * (size_t)nsub * 8)) return bail(rc);
    cudaMemset(out_deg.p, 0, (size_t)V * 8);
    cudaMemset(nrm.p, 0, (size_t)V * 8);
    cudaMemset(ssum.p, 0, 8);
    cudaMemset(cnt.p, 0, 2 * (size_t)nsub * 8);
    k_synth_degrees<<<grid, 256>>>(s, (double*)out_deg.p, (double*)nrm.p);
    k_synth_psrc<<<grid, 256>>>(V, (const double*)out_deg.p, (const double*)nrm.p, cpp, (double*)psrc.p, (double*)ssum.p);
    k_synth_count<<<grid, 256>>>(s, gm, (unsigned long long*)cnt.p);
    g_launches += 3;
    if (cudaGetLastError() != cudaSuccess) return bail(fail(SMORE_E_CUDA, "synthetic graph kernels failed to launch"));
    double src_sum = 0;
    std::vector<unsigned long long> hcnt((size_t)nsub);
    if (cudaMemcpy(&src_sum, ssum.p, 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(hcnt.data(), cnt.p, (size_t)nsub * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
        return bail(fail(SMORE_E_CUDA, "synthetic graph: %s", cudaGetErrorString(cudaGetLastError())));
    if (!(src_sum > 0)) return bail(fail(SMORE_E_INVALID, "the synthetic graph has no edges"));
    std::vector<int64_t> off((size_t)nsub + 1, 0);
    for (int q = 0; q < nsub; ++q) off[(size_t)q + 1] = off[(size_t)q] + (int64_t)hcnt[(size_t)q];
    const int64_t ne = off[(size_t)nsub];
    if (ne == 0) return bail(fail(SMORE_E_INVALID, "rank %d owns no edge target", rank));
    if (ne >= (1ll << 32)) return bail(fail(SMORE_E_UNSUPPORTED, "rank %d owns %lld entries; edge tables use 32-bit draws", rank, (long long)ne));
    DevBuf pe;
    if (int rc = pe.ensure((size_t)ne * 8)) return bail(rc);
    if (cudaMalloc((void**)&g->d_eat, (size_t)ne * sizeof(uint2)) != cudaSuccess || cudaMalloc((void**)&g->d_esrc, (size_t)ne * 4) != cudaSuccess ||
        cudaMalloc((void**)&g->d_edst, (size_t)ne * 4) != cudaSuccess)
        return bail(fail(SMORE_E_NOMEM, "cudaMalloc of the block tables (%lld entries) failed", (long long)ne));
    {
        std::vector<unsigned long long> fill((size_t)nsub);
        for (int q = 0; q < nsub; ++q) fill[(size_t)q] = (unsigned long long)off[(size_t)q];
        unsigned long long* d_fill = (unsigned long long*)cnt.p + nsub;
        cudaMemcpy(d_fill, fill.data(), (size_t)nsub * 8, cudaMemcpyHostToDevice);
        k_synth_fill<<<grid, 256>>>(s, gm, (const double*)psrc.p, (const double*)nrm.p, (const double*)out_deg.p, cpp, 1.0 / src_sum,
                                    d_fill, (double*)pe.p, g->d_esrc, g->d_edst);
        g_launches++;
    }
    AliasScratchDev scr;
    g->blk_mass.assign((size_t)nsub, 0.0);
    double own = 0;
    for (int q = 0; q < nsub; ++q) {
        const int64_t n = off[(size_t)q + 1] - off[(size_t)q];
        if (n == 0) continue;
        double mass = 0;
        if (int rc = alias_build_device((const double*)pe.p + off[(size_t)q], n, g->d_eat + off[(size_t)q], scr, &mass)) return bail(rc);
        g->blk_mass[(size_t)q] = mass;  // the entry probabilities sum to 1 over the whole graph
        own += mass;
    }
    // shard-local negative table and the per-sub-part vertex tables
    if (int rc = wtmp.ensure((size_t)std::max(nl, sub_cap) * 8)) return bail(rc);
    if (cudaMalloc((void**)&g->d_nat, (size_t)nl * sizeof(uint2)) != cudaSuccess) return bail(fail(SMORE_E_NOMEM, "cudaMalloc failed"));
    k_synth_neg_weights<<<grid, 256>>>(nl, world, rank, (const double*)out_deg.p, (double*)wtmp.p);
    g_launches++;
    if (int rc = alias_build_device((const double*)wtmp.p, nl, g->d_nat, scr)) return bail(rc);
    g->vsub_off.assign((size_t)nsub + 1, 0);
    g->sub_rows.assign((size_t)nsub, 0);
    for (int q = 0; q < nsub; ++q) {
        const int64_t rows = (V - (q >> 1) + world - 1) / world;
        g->sub_rows[(size_t)q] = (q & 1) ? std::max<int64_t>(0, rows - sub_cap) : std::min(rows, sub_cap);
        g->vsub_off[(size_t)q + 1] = g->vsub_off[(size_t)q] + g->sub_rows[(size_t)q];
    }
    if (cudaMalloc((void**)&g->d_vsub, (size_t)g->vsub_off[(size_t)nsub] * sizeof(uint2)) != cudaSuccess) return bail(fail(SMORE_E_NOMEM, "cudaMalloc failed"));
    for (int q = 0; q < nsub; ++q) {
        const int64_t n = g->sub_rows[(size_t)q];
        if (n == 0) continue;
        k_synth_sub_weights<<<grid, 256>>>(n, q, shift, sub_cap, (const double*)psrc.p, (double*)wtmp.p);
        g_launches++;
        if (int rc = alias_build_device((const double*)wtmp.p, n, g->d_vsub + g->vsub_off[(size_t)q], scr)) return bail(rc);
    }
    if (int rc = upload_lut(g)) return bail(rc);
    if (cudaDeviceSynchronize() != cudaSuccess) return bail(fail(SMORE_E_CUDA, "synthetic graph: %s", cudaGetErrorString(cudaGetLastError())));
    g->n_edge_local = ne;
    g->rank = rank;
    g->world = world;
    g->shift = shift;
    g->n_local = nl;
    g->n_neg = nl;
    g->rotating = true;
    g->nsub = nsub;
    g->sub_cap = sub_cap;
    g->blk_off = off;
    g->src_mass_frac = own;
    *out = g;
    return SMORE_OK;
}

}  // extern "C"
