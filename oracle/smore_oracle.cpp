// TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's sampled-SGD hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library; the product (smore_b200/) never imports, links or executes anything under oracle/.
//
// What it restates (file:line are into /root/reference):
//   * C++ tree ("cpp" semantics): src/proNet.cpp  AliasMethod :544-620, BuildAliasMethod :410-542,
//     samplers :623-683, RandomWalk :704-724, SkipGrams :769-809, ScaleSkipGrams :928-987,
//     InitSigmoid/fastSigmoid :52-71, Opt_SigmoidSGD :1312-1330, Opt_BPRSGD :1053-1068, Opt_FBPRSGD :1014-1031,
//     UpdatePair :1784-1809, UpdateBPRPair :1406-1455, UpdateWARPPair :1353-1403, UpdateFBPRPair :1458-1515;
//     Train loops src/model/LINE.cpp:100-195, BPR.cpp:55-107, WARP.cpp:55-107, HBPR.cpp:63-130,
//     DeepWalk.cpp:98-155, Walklets.cpp:6-64.
//   * Go tree ("go" semantics): pkg/pronet/alias.go:10-106, pkg/pronet/pronet.go:90-109,182-333,
//     pkg/pronet/optimizer.go:8-117, internal/models/line/line.go:73-206, bpr/bpr.go:61-131,
//     deepwalk/deepwalk.go:61-141.
//
// Randomness: the reference's RNGs (thread_local mt19937 / Go math/rand) are NOT replayed. Every draw pops one
// 32-bit word k of a Philox4x32-10 stream exactly as oracle/ref_shim.cpp does for the compiled C++ reference:
//   probability draw = k * 2^-32 ; index draw over n = floor(k*n / 2^32) ; libc rand() = (shuffle-stream k) >> 1.
// For the Go tree the same mapping stands in for rng.Float64() / rng.Intn(n) / rand.Int63n(n).
//
// Pinning: the reference ships no tests or golden vectors, so parity is pinned against outputs of the
// reference itself: tests/test_oracle_vs_ref.py compares every C++-semantics function here with oracle/_ref/
// libsmore_ref.so (the unmodified reference compiled with the shim) and tests/golden/* holds vectors generated from it.
// The Go tree cannot be compiled here (no Go toolchain): Go-semantics functions are "parity unpinned" except where
// they coincide with the C++ algorithm (alias build at power 0.75, sigmoid LUT).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "oracle_philox.h"

namespace {

constexpr int MONITOR = 10000;            // src/proNet.h:32, pronet.go:15
constexpr double POWER_SAMPLE = 0.75;     // src/proNet.h:33
constexpr int SIGMOID_TABLE_SIZE = 1000;  // src/proNet.h:35
constexpr double MAX_SIGMOID = 8.0;       // src/proNet.h:36

enum Semantics { SEM_CPP = 0, SEM_GO = 1 };
enum NegMethod { NEG_DEGREES = 0, NEG_IN_DEGREES = 1, NEG_NO_DEGREES = 2 };

struct Alias {
    std::vector<double> prob;
    std::vector<int64_t> alias;
};

// ---- draw stream (same mapping as ref_shim.cpp) ------------------------------------------------------------
struct Draws {
    oracle_stream s;
    oracle_stream shuffle;
    Draws(uint64_t seed, uint64_t stream) {
        oracle_stream_init(&s, seed, stream, 0);
        oracle_stream_init(&shuffle, seed, ORACLE_SHUFFLE_STREAM, 0);
    }
    double prob() { return (double)oracle_stream_next(&s) * (1.0 / 4294967296.0); }
    // random_gen(0, n) as the shim returns it, truncated to long by the caller
    int64_t index(int64_t n) {
        uint32_t k = oracle_stream_next(&s);
        if (n == 1) return (int64_t)((double)k * (1.0 / 4294967296.0));  // == 0
        return (int64_t)(((uint64_t)k * (uint64_t)n) >> 32);
    }
    int libc_rand() { return (int)(oracle_stream_next(&shuffle) >> 1); }
    int64_t shuffle_index(int64_t n) { return (int64_t)(((uint64_t)oracle_stream_next(&shuffle) * (uint64_t)n) >> 32); }
};

// ---- alias construction ------------------------------------------------------------------------------------
// C++: src/proNet.cpp:544-620. The `power` argument is ignored: POWER_SAMPLE is always applied (:558,:564).
Alias alias_cpp(const std::vector<double>& dist) {
    size_t n = dist.size();
    Alias t;
    t.prob.assign(n, 0.0);
    t.alias.assign(n, -1);
    double sum = 0;
    for (size_t i = 0; i < n; ++i) sum += pow(dist[i], POWER_SAMPLE);
    double norm = n / sum;
    std::vector<double> np(n);
    for (size_t i = 0; i < n; ++i) np[i] = pow(dist[i], POWER_SAMPLE) * norm;
    std::vector<int64_t> small, large;
    for (size_t i = 0; i < n; ++i) (np[i] < 1 ? small : large).push_back((int64_t)i);
    while (!small.empty() && !large.empty()) {
        int64_t s = small.back(); small.pop_back();
        int64_t l = large.back(); large.pop_back();
        t.alias[s] = l;
        t.prob[s] = np[s];
        np[l] = np[l] + np[s] - 1;
        (np[l] < 1 ? small : large).push_back(l);
    }
    while (!large.empty()) { t.prob[large.back()] = 1.0; large.pop_back(); }
    while (!small.empty()) { t.prob[small.back()] = 1.0; small.pop_back(); }
    return t;
}

// Go: pkg/pronet/alias.go:10-90. power honoured, zero weights -> 0 (no Pow), all-zero -> uniform, alias=self on prob 1.
Alias alias_go(const std::vector<double>& dist, double power) {
    size_t n = dist.size();
    Alias t;
    t.prob.assign(n, 0.0);
    t.alias.assign(n, 0);
    if (n == 0) return t;
    double sum = 0.0;
    std::vector<double> norm(n);
    for (size_t i = 0; i < n; ++i) {
        norm[i] = dist[i] > 0 ? pow(dist[i], power) : 0.0;
        sum += norm[i];
    }
    if (sum == 0) {
        for (size_t i = 0; i < n; ++i) { t.prob[i] = 1.0; t.alias[i] = (int64_t)i; }
        return t;
    }
    for (size_t i = 0; i < n; ++i) norm[i] = norm[i] * (double)n / sum;
    std::vector<int64_t> small, large;
    for (size_t i = 0; i < n; ++i) (norm[i] < 1.0 ? small : large).push_back((int64_t)i);
    while (!small.empty() && !large.empty()) {
        int64_t l = small.back(); small.pop_back();
        int64_t g = large.back(); large.pop_back();
        t.prob[l] = norm[l];
        t.alias[l] = g;
        norm[g] = norm[g] + norm[l] - 1.0;
        (norm[g] < 1.0 ? small : large).push_back(g);
    }
    while (!large.empty()) { int64_t g = large.back(); large.pop_back(); t.prob[g] = 1.0; t.alias[g] = g; }
    while (!small.empty()) { int64_t l = small.back(); small.pop_back(); t.prob[l] = 1.0; t.alias[l] = l; }
    return t;
}

// ---- graph -------------------------------------------------------------------------------------------------
struct Graph {
    int sem;
    int64_t V, E;
    int64_t max_line;  // C++: CSR entries (lines, doubled if undirected :230-231); Go: edge lines (pronet.go:164)
    std::vector<int64_t> off;
    std::vector<int32_t> col;
    std::vector<double> w;
    std::vector<double> out_deg, in_deg;
    Alias vertex_at, negative_at;
    Alias context_at;           // C++ only: per-vertex sub-tables in CSR order, alias translated to vid (:519-536)
    std::vector<int32_t> field;  // HOP-Rec: field[vid].fields[0] (src/proNet.cpp:330-408); default 0
    double sigmoid[SIGMOID_TABLE_SIZE + 1];

    void init_sigmoid() {  // src/proNet.cpp:52-60 (incl. the one-past-the-end entry), pronet.go:90-95
        for (int i = 0; i != SIGMOID_TABLE_SIZE + 1; i++) {
            double x = i * 2.0 * MAX_SIGMOID / SIGMOID_TABLE_SIZE - MAX_SIGMOID;
            sigmoid[i] = 1.0 / (1.0 + exp(-x));
        }
    }
    double fast_sigmoid(double x) const {  // src/proNet.cpp:62-71 ; pronet.go:98-109 (its idx clamp is unreachable)
        if (x < -MAX_SIGMOID) return 0.0;
        if (x > MAX_SIGMOID) return 1.0;
        return sigmoid[int((x + MAX_SIGMOID) * SIGMOID_TABLE_SIZE / MAX_SIGMOID / 2)];
    }

    void build(int neg_method) {
        out_deg.assign(V, 0.0);
        in_deg.assign(V, 0.0);
        // src/proNet.cpp:423-446 ; pronet.go:198-212 (same accumulation order: CSR order)
        for (int64_t v = 0; v < V; ++v)
            for (int64_t e = off[v]; e < off[v + 1]; ++e) out_deg[v] += w[e];
        for (int64_t e = 0; e < E; ++e) in_deg[col[e]] += w[e];
        field.assign(V, 0);
        init_sigmoid();

        std::vector<double> dist(V);
        if (sem == SEM_CPP) {
            for (int64_t v = 0; v < V; ++v) dist[v] = out_deg[v];  // vertex_method "out_degrees" :458-464
            vertex_at = alias_cpp(dist);
            for (int64_t v = 0; v < V; ++v) {  // :486-509
                if (neg_method == NEG_DEGREES) dist[v] = in_deg[v] + out_deg[v];
                else if (neg_method == NEG_IN_DEGREES) dist[v] = in_deg[v];
                else dist[v] = in_deg[v] == 0 ? 0 : 1;
            }
            negative_at = alias_cpp(dist);
            context_at.prob.resize(E);
            context_at.alias.resize(E);
            std::vector<double> sub;
            for (int64_t v = 0; v < V; ++v) {  // :519-536
                int64_t o = off[v], b = off[v + 1] - off[v];
                sub.assign(w.begin() + o, w.begin() + o + b);
                Alias a = alias_cpp(sub);
                for (int64_t i = 0; i < b; ++i) {
                    context_at.prob[o + i] = a.prob[i];
                    context_at.alias[o + i] = a.alias[i] != -1 ? col[o + a.alias[i]] : -1;
                }
            }
        } else {
            for (int64_t v = 0; v < V; ++v) dist[v] = out_deg[v];  // pronet.go:224-230
            vertex_at = alias_go(dist, 1.0);
            for (int64_t v = 0; v < V; ++v) dist[v] = in_deg[v] + out_deg[v];  // pronet.go:242-249
            negative_at = alias_go(dist, POWER_SAMPLE);
        }
    }

    // ---- samplers ------------------------------------------------------------------------------------------
    int64_t source_sample(Draws& d) const {
        if (sem == SEM_CPP) {  // src/proNet.cpp:647-657 : p first, then index
            double p = d.prob();
            int64_t v = d.index(V);
            return p < vertex_at.prob[v] ? v : vertex_at.alias[v];
        }
        int64_t i = d.index(V);  // alias.go:93-106 : Intn then Float64
        double r = d.prob();
        return r < vertex_at.prob[i] ? i : vertex_at.alias[i];
    }
    int64_t negative_sample(Draws& d) const {
        int64_t v = d.index(V);  // src/proNet.cpp:623-633 : index first, then p (Go identical order)
        double p = d.prob();
        return p < negative_at.prob[v] ? v : negative_at.alias[v];
    }
    int64_t target_sample(int64_t vid, Draws& d) const {
        int64_t branch = off[vid + 1] - off[vid];
        if (branch == 0) return -1;
        if (sem == SEM_CPP) {  // src/proNet.cpp:671-683
            double p = d.prob();
            int64_t j = d.index(branch) + off[vid];
            return p < context_at.prob[j] ? col[j] : context_at.alias[j];
        }
        // pronet.go:257-284 : O(deg) CDF scan, `r <= cum`
        double total = 0.0;
        for (int64_t e = off[vid]; e < off[vid + 1]; ++e) total += w[e];
        double r = d.prob() * total;
        double cum = 0.0;
        for (int64_t e = off[vid]; e < off[vid + 1]; ++e) {
            cum += w[e];
            if (r <= cum) return col[e];
        }
        return col[off[vid + 1] - 1];
    }

    // ---- walks -----------------------------------------------------------------------------------------------
    void random_walk(int64_t start, int steps, Draws& d, std::vector<int64_t>& walk) const {
        walk.clear();
        walk.push_back(start);
        int64_t next = start;
        if (sem == SEM_CPP) {  // src/proNet.cpp:704-724 : dead end -> teleport to start
            for (int s = 0; s < steps; ++s) {
                if (off[next + 1] - off[next] == 0) {
                    if (next == start) return;
                    next = start;
                }
                next = target_sample(next, d);
                walk.push_back(next);
            }
        } else {  // pronet.go:292-307 : dead end -> stop
            for (int s = 0; s < steps; ++s) {
                int64_t n = target_sample(next, d);
                if (n == -1) break;
                walk.push_back(n);
                next = n;
            }
        }
    }
    // node2vec (Go tree only): internal/models/node2vec/node2vec.go:113-173. areNeighbors is a linear scan of prev's adjacency.
    bool are_neighbors(int64_t a, int64_t b) const {
        for (int64_t e = off[a]; e < off[a + 1]; ++e)
            if (col[e] == b) return true;
        return false;
    }
    int64_t biased_target_sample(int64_t prev, int64_t cur, double p, double q, Draws& d, std::vector<double>& bw) const {
        int64_t branch = off[cur + 1] - off[cur];
        if (branch == 0) return -1;
        bw.resize((size_t)branch);
        double total = 0.0;
        for (int64_t i = 0; i < branch; ++i) {
            int64_t nb = col[off[cur] + i];
            double bias;
            if (nb == prev) bias = 1.0 / p;
            else if (are_neighbors(prev, nb)) bias = 1.0;
            else bias = 1.0 / q;
            bw[(size_t)i] = w[off[cur] + i] * bias;
            total += bw[(size_t)i];
        }
        if (total == 0) return col[off[cur] + d.index(branch)];
        double r = d.prob() * total;
        double cum = 0.0;
        for (int64_t i = 0; i < branch; ++i) {
            cum += bw[(size_t)i];
            if (r <= cum) return col[off[cur] + i];
        }
        return col[off[cur + 1] - 1];
    }
    void biased_random_walk(int64_t start, int steps, double p, double q, Draws& d, std::vector<int64_t>& walk,
                            std::vector<double>& bw) const {  // node2vec.go:82-110
        walk.clear();
        walk.push_back(start);
        if (steps == 0) return;
        int64_t first = target_sample(start, d);
        if (first == -1) return;
        walk.push_back(first);
        for (int i = 1; i < steps; ++i) {
            int64_t nxt = biased_target_sample(walk[walk.size() - 2], walk.back(), p, q, d, bw);
            if (nxt == -1) break;
            walk.push_back(nxt);
        }
    }
    void skip_grams(const std::vector<int64_t>& walk, int window, Draws& d, std::vector<int64_t>& pv,
                    std::vector<int64_t>& pc) const {
        pv.clear(); pc.clear();
        int length = (int)walk.size();
        if (sem == SEM_CPP) {  // src/proNet.cpp:769-809 : one draw per centre (:783)
            for (int i = 0; i < length; ++i) {
                int reduce = (int)d.index(window) + 1;
                int left = std::max(i - reduce, 0);
                int right = i + reduce;
                if (right >= length) right = length - 1;
                for (int j = left; j <= right; ++j) {
                    if (i == j) continue;
                    pv.push_back(walk[i]); pc.push_back(walk[j]);
                }
            }
        } else {  // pronet.go:310-333 : full fixed window
            for (int i = 0; i < length; ++i) {
                int start = std::max(i - window, 0);
                int end = std::min(i + window + 1, length);
                for (int j = start; j < end; ++j)
                    if (i != j) { pv.push_back(walk[i]); pc.push_back(walk[j]); }
            }
        }
    }
    // src/proNet.cpp:928-987 (negative_samples == 0 as Walklets.cpp:47 calls it)
    void scale_skip_grams(const std::vector<int64_t>& walk, int wmin, int wmax, std::vector<int64_t>& pv,
                          std::vector<int64_t>& pc) const {
        pv.clear(); pc.clear();
        int length = (int)walk.size();
        for (int i = 0; i < length; ++i) {
            int left = std::max(i - wmax, 0);
            int right = std::max(i - wmin, 0);
            for (int j = left; j <= right; ++j) {
                if (i == j) continue;
                pv.push_back(walk[i]); pc.push_back(walk[j]);
            }
            left = i + wmin; if (left >= length) left = length - 1;
            right = i + wmax; if (right >= length) right = length - 1;
            for (int j = left; j <= right; ++j) {
                if (i == j) continue;
                pv.push_back(walk[i]); pc.push_back(walk[j]);
            }
        }
    }

    // ---- SGD primitives (operate on raw rows; rows may alias exactly as in the reference) ---------------------
    // src/proNet.cpp:1312-1330
    void opt_sigmoid_sgd(const double* wv, const double* wc, double label, int dim, double alpha, double* loss_v,
                         double* loss_c) const {
        double f = 0;
        for (int d = 0; d < dim; ++d) f += wv[d] * wc[d];
        f = fast_sigmoid(f);
        double g = (label - f) * alpha;
        for (int d = 0; d < dim; ++d) loss_v[d] += g * wc[d];
        for (int d = 0; d < dim; ++d) loss_c[d] += g * wv[d];
    }
    // src/proNet.cpp:1784-1809
    void update_pair_cpp(double* Wv, double* Wc, int64_t vertex, int64_t context, int dim, int K, double alpha,
                         Draws& d, std::vector<double>& back_err) const {
        back_err.assign(dim, 0.0);
        double* rv = Wv + vertex * dim;
        double* rc = Wc + context * dim;
        opt_sigmoid_sgd(rv, rc, 1.0, dim, alpha, back_err.data(), rc);
        for (int n = 0; n != K; ++n) {
            int64_t c = negative_sample(d);
            rc = Wc + c * dim;
            opt_sigmoid_sgd(rv, rc, 0.0, dim, alpha, back_err.data(), rc);
        }
        for (int k = 0; k < dim; ++k) rv[k] += back_err[k];
    }
    // src/proNet.cpp:1332-1351 (HPE): the regularised variant; alpha multiplies the whole bracket
    void opt_sigmoid_reg_sgd(const double* wv, const double* wc, double label, int dim, double alpha, double reg,
                             double* loss_v, double* loss_c) const {
        double f = 0;
        for (int d = 0; d < dim; ++d) f += wv[d] * wc[d];
        f = fast_sigmoid(f);
        double g = (label - f);
        for (int d = 0; d < dim; ++d) loss_v[d] += alpha * (g * wc[d] - reg * wv[d]);
        for (int d = 0; d < dim; ++d) loss_c[d] += alpha * (g * wv[d] - reg * wc[d]);
    }
    // src/proNet.cpp:3018-3054: the context walks on (TargetSample) for walk_steps steps, each step one positive and K
    // negatives against the SAME vertex row, which is updated at the end of every step
    void update_community_cpp(double* Wv, double* Wc, int64_t vertex, int64_t context, int dim, double reg, int walk_steps,
                              int K, double alpha, Draws& d, std::vector<double>& back_err) const {
        double* rv = Wv + vertex * dim;
        for (int s = 0; s < walk_steps; s++) {
            if (s != 0) {
                context = target_sample(context, d);
                if (context == -1) break;
            }
            back_err.assign(dim, 0.0);
            double* rc = Wc + context * dim;
            opt_sigmoid_reg_sgd(rv, rc, 1.0, dim, alpha, reg, back_err.data(), rc);
            for (int n = 0; n != K; ++n) {
                int64_t c = negative_sample(d);
                rc = Wc + c * dim;
                opt_sigmoid_reg_sgd(rv, rc, 0.0, dim, alpha, reg, back_err.data(), rc);
            }
            for (int k = 0; k < dim; ++k) rv[k] += back_err[k];
        }
    }
    // src/proNet.cpp:1070-1098 (Skew-OPT): g = (f - xi)/omega, gated at g > 2, clamped at -2, then the eta-th power chain
    int opt_sbpr_sgd(const double* wv, const double* wc, int dim, double xi, double omega, int eta, double alpha,
                     double* loss_v, double* loss_c) const {
        double f = 0, g = 0, g_in_sigmoid = 1, g_chain_diff = 1;
        for (int d = 0; d < dim; ++d) f += wv[d] * wc[d];
        g = (f - xi) / omega;
        if (g > 2.0) return 0;
        if (g < -2.0) g = -2.0;
        for (int i = 0; i < eta; i++) g_in_sigmoid *= g;
        g_chain_diff = g_in_sigmoid / g;
        g = fast_sigmoid(-1 * g_in_sigmoid) * g_chain_diff / omega;
        g *= alpha;
        for (int d = 0; d < dim; ++d) loss_v[d] += g * wc[d];
        for (int d = 0; d < dim; ++d) loss_c[d] += g * wv[d];
        return 1;
    }
    // src/proNet.cpp:1517-1566: 16 rounds, every negative drawn here; accepted rounds decay and move both item rows,
    // the user row moves once by the averaged error (one shared table: rows may alias)
    void update_sbpr_pair_cpp(double* W, int64_t vertex, int64_t ci, int dim, double xi, double omega, int eta, double alpha,
                              Draws& d, std::vector<double>& verr, std::vector<double>& cerr, std::vector<double>& cvec) const {
        verr.assign(dim, 0.0); cerr.assign(dim, 0.0); cvec.assign(dim, 0.0);
        double* rv = W + vertex * dim;
        double* ri = W + ci * dim;
        int update = 0;
        for (int n = 0; n < 16; n++) {
            int64_t cj = negative_sample(d);
            double* rj = W + cj * dim;
            for (int k = 0; k < dim; k++) {
                cerr[k] = 0.0;
                cvec[k] = ri[k] - rj[k];
            }
            if (opt_sbpr_sgd(rv, cvec.data(), dim, xi, omega, eta, alpha, verr.data(), cerr.data()) != 0) {
                for (int k = 0; k < dim; k++) {
                    ri[k] -= alpha * 0.01 * ri[k];
                    rj[k] -= alpha * 0.01 * rj[k];
                    ri[k] += cerr[k];
                    rj[k] -= cerr[k];
                }
                update += 1.0;
            }
        }
        if (update != 0)
            for (int k = 0; k < dim; k++) {
                rv[k] -= alpha * 0.01 * rv[k];
                rv[k] += verr[k] / update;
            }
    }
    // src/proNet.cpp:991-1012 (MF): linear prediction, no sigmoid
    void opt_sgd(const double* wv, const double* wc, double label, int dim, double alpha, double reg, double* loss_v,
                 double* loss_c) const {
        double f = 0;
        for (int d = 0; d < dim; ++d) f += wv[d] * wc[d];
        double g = (label - f);
        for (int d = 0; d < dim; ++d) loss_v[d] += alpha * (g * wc[d] - reg * wv[d]);
        for (int d = 0; d < dim; ++d) loss_c[d] += alpha * (g * wv[d] - reg * wc[d]);
    }
    // src/proNet.cpp:2591-2614: labels +1 / -1; MF::Train passes ONE table in both roles, so rows may alias
    void update_factorized_pair_cpp(double* Wv, double* Wc, int64_t vertex, int64_t context, int dim, double reg, int K,
                                    double alpha, Draws& d, std::vector<double>& back_err) const {
        back_err.assign(dim, 0.0);
        double* rv = Wv + vertex * dim;
        double* rc = Wc + context * dim;
        opt_sgd(rv, rc, 1.0, dim, alpha, reg, back_err.data(), rc);
        for (int n = 0; n != K; ++n) {
            int64_t c = negative_sample(d);
            rc = Wc + c * dim;
            opt_sgd(rv, rc, -1.0, dim, alpha, reg, back_err.data(), rc);
        }
        for (int k = 0; k < dim; ++k) rv[k] += back_err[k];
    }
    // optimizer.go:21-58 + sgdUpdate :61-84
    void update_pair_go(double* Wv, double* Wc, int64_t vertex, int64_t context, int dim, int K, double alpha, Draws& d,
                        std::vector<double>& vgrad, std::vector<double>& cgrad, std::vector<double>& ngrad) const {
        vgrad.assign(dim, 0.0);
        cgrad.assign(dim, 0.0);
        auto sgd = [&](const double* ve, const double* ce, double label, double* vg, double* cg) {
            double score = 0.0;
            for (int k = 0; k < dim; ++k) score += ve[k] * ce[k];
            double pred = fast_sigmoid(score);
            double grad = alpha * (label - pred);
            for (int k = 0; k < dim; ++k) {
                vg[k] += grad * ce[k];
                cg[k] += grad * ve[k];
            }
        };
        double* rv = Wv + vertex * dim;
        sgd(rv, Wc + context * dim, 1.0, vgrad.data(), cgrad.data());
        for (int i = 0; i < K; ++i) {
            int64_t neg = negative_sample(d);
            if (neg == context) continue;
            ngrad.assign(dim, 0.0);
            double* rn = Wc + neg * dim;
            sgd(rv, rn, 0.0, vgrad.data(), ngrad.data());
            for (int k = 0; k < dim; ++k) rn[k] += ngrad[k];
        }
        double* rc = Wc + context * dim;
        for (int k = 0; k < dim; ++k) {
            rv[k] += vgrad[k];
            rc[k] += cgrad[k];
        }
    }
    // line.go:153-200
    void update_first_order_go(double* W, int64_t source, int64_t target, int dim, int K, double alpha, Draws& d,
                               std::vector<double>& vgrad, std::vector<double>& cgrad) const {
        vgrad.assign(dim, 0.0);
        cgrad.assign(dim, 0.0);
        double* rs = W + source * dim;
        double* rt = W + target * dim;
        double score = 0.0;
        for (int k = 0; k < dim; ++k) score += rs[k] * rt[k];
        double grad = alpha * (1.0 - fast_sigmoid(score));
        for (int k = 0; k < dim; ++k) {
            vgrad[k] = grad * rt[k];
            cgrad[k] = grad * rs[k];
        }
        for (int i = 0; i < K; ++i) {
            int64_t neg = negative_sample(d);
            if (neg == target || neg == source) continue;
            double* rn = W + neg * dim;
            double sc = 0.0;
            for (int k = 0; k < dim; ++k) sc += rs[k] * rn[k];
            double g = alpha * (0.0 - fast_sigmoid(sc));
            for (int k = 0; k < dim; ++k) {
                vgrad[k] += g * rn[k];
                rn[k] += g * rs[k];
            }
        }
        for (int k = 0; k < dim; ++k) {
            rs[k] += vgrad[k];
            rt[k] += cgrad[k];
        }
    }
    // src/proNet.cpp:1053-1068
    void opt_bpr_sgd(const double* wv, const double* wc, int dim, double alpha, double* loss_v, double* loss_c) const {
        double f = 0;
        for (int d = 0; d < dim; ++d) f += wv[d] * wc[d];
        double g = fast_sigmoid(0.0 - f) * alpha;
        for (int d = 0; d < dim; ++d) loss_v[d] += g * wc[d];
        for (int d = 0; d < dim; ++d) loss_c[d] += g * wv[d];
    }
    // src/proNet.cpp:1014-1031
    int opt_fbpr_sgd(const double* wv, const double* wc, int dim, double alpha, double* loss_v, double* loss_c,
                     double e) const {
        double f = 0;
        for (int d = 0; d < dim; ++d) f += wv[d] * wc[d];
        if (f > e) return 0;
        double g = fast_sigmoid(0.0 - f) * alpha;
        for (int d = 0; d < dim; ++d) loss_v[d] += g * wc[d];
        for (int d = 0; d < dim; ++d) loss_c[d] += g * wv[d];
        return 1;
    }
    // src/proNet.cpp:1406-1455 (one shared table; `reg` unused)
    void update_bpr_pair_cpp(double* W, int64_t vertex, int64_t ci, int64_t cj, int dim, double alpha, Draws& d,
                             std::vector<double>& verr, std::vector<double>& cerr, std::vector<double>& cvec) const {
        verr.assign(dim, 0.0); cerr.assign(dim, 0.0); cvec.assign(dim, 0.0);
        double* rv = W + vertex * dim;
        double* ri = W + ci * dim;
        for (int n = 0; n < 5; n++) {
            if (n != 0) cj = negative_sample(d);
            double* rj = W + cj * dim;
            for (int k = 0; k < dim; k++) {
                cerr[k] = 0.0;
                cvec[k] = ri[k] - rj[k];
            }
            opt_bpr_sgd(rv, cvec.data(), dim, alpha, verr.data(), cerr.data());
            for (int k = 0; k < dim; k++) {
                ri[k] -= alpha * 0.0025 * ri[k];
                rj[k] -= alpha * 0.0025 * rj[k];
                ri[k] += cerr[k];
                rj[k] -= cerr[k];
            }
        }
        for (int k = 0; k < dim; k++) {
            rv[k] -= alpha * 0.025 * rv[k];
            rv[k] += verr[k];
        }
    }
    // src/proNet.cpp:1353-1403 ; returns number of negatives scanned (diagnostic for the byte model)
    int update_warp_pair_cpp(double* W, int64_t vertex, int64_t ci, int64_t cj, int dim, double alpha, Draws& d,
                             std::vector<double>& verr, std::vector<double>& cerr, std::vector<double>& cvec) const {
        verr.assign(dim, 0.0); cerr.assign(dim, 0.0); cvec.assign(dim, 0.0);
        double* rv = W + vertex * dim;
        double* ri = W + ci * dim;
        int tries = 0;
        for (int n = 0; n < 32; n++) {
            double f = 0.0;
            if (n != 0) cj = negative_sample(d);
            double* rj = W + cj * dim;
            ++tries;
            for (int k = 0; k < dim; k++) {
                cvec[k] = ri[k] - rj[k];
                f += rv[k] * cvec[k];
            }
            if (f < 1.0) {
                opt_bpr_sgd(rv, cvec.data(), dim, alpha, verr.data(), cerr.data());
                for (int k = 0; k < dim; k++) {
                    ri[k] -= alpha * 0.0025 * ri[k];
                    rj[k] -= alpha * 0.0025 * rj[k];
                    rv[k] -= alpha * 0.0025 * rv[k];
                    ri[k] += cerr[k];
                    rj[k] -= cerr[k];
                    rv[k] += verr[k];
                }
                break;
            }
        }
        return tries;
    }
    // src/proNet.cpp:1458-1515
    void update_fbpr_pair_cpp(double* W, int64_t vertex, int64_t ci, int64_t cj, int dim, double alpha, double margin,
                              Draws& d, std::vector<double>& verr, std::vector<double>& cerr,
                              std::vector<double>& cvec) const {
        verr.assign(dim, 0.0); cerr.assign(dim, 0.0); cvec.assign(dim, 0.0);
        double* rv = W + vertex * dim;
        double* ri = W + ci * dim;
        double up = 0.0;
        for (int w = 0; w < 5; w++) {
            if (w != 0) {
                cj = d.index(V);
                while (field[ci] != field[cj]) cj = d.index(V);
            }
            double* rj = W + cj * dim;
            for (int k = 0; k < dim; k++) {
                cerr[k] = 0.0;
                cvec[k] = ri[k] - rj[k];
            }
            if (opt_fbpr_sgd(rv, cvec.data(), dim, alpha, verr.data(), cerr.data(), margin)) {
                up++;
                for (int k = 0; k < dim; k++) {
                    ri[k] -= alpha * 0.0025 * ri[k];
                    rj[k] -= alpha * 0.0025 * rj[k];
                    ri[k] += cerr[k];
                    rj[k] -= cerr[k];
                }
            }
        }
        if (up)
            for (int k = 0; k < dim; k++) {
                rv[k] -= alpha * 0.025 * rv[k];
                rv[k] += verr[k] / up;
            }
    }
    // optimizer.go:87-117
    void update_bpr_pair_go(double* Wv, double* Wc, int64_t vertex, int64_t pos, int64_t neg, int dim, double alpha,
                            double lambda) const {
        double* rv = Wv + vertex * dim;
        double* rp = Wc + pos * dim;
        double* rn = Wc + neg * dim;
        double ps = 0.0, ns = 0.0;
        for (int k = 0; k < dim; ++k) {
            ps += rv[k] * rp[k];
            ns += rv[k] * rn[k];
        }
        double diff = ns - ps;
        double gc = alpha * fast_sigmoid(diff);
        for (int k = 0; k < dim; ++k) {
            double vg = gc * (rp[k] - rn[k]);
            double pg = gc * rv[k];
            double ng = -gc * rv[k];
            rv[k] += vg - lambda * alpha * rv[k];
            rp[k] += pg - lambda * alpha * rp[k];
            rn[k] += ng - lambda * alpha * rn[k];
        }
    }
};

// LR tick shared by LINE/BPR/WARP/HBPR C++ loops (LINE.cpp:177-187): note alpha uses current_sample BEFORE the bump.
struct CppSchedule {
    double alpha, alpha_min, cur;
    unsigned long long total, current_sample = 0;
    CppSchedule(double a, unsigned long long t) : alpha(a), alpha_min(a * 0.0001), cur(a), total(t) {}
    void tick(unsigned long long count) {
        if (count % MONITOR == 0) {
            cur = alpha * (1.0 - (double)(current_sample) / total);
            current_sample += MONITOR;
            if (cur < alpha_min) cur = alpha_min;
        }
    }
};

}  // namespace

extern "C" {

// ---- stream / philox -------------------------------------------------------------------------------------
void orc_philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { oracle_philox4x32_10(ctr, key, out); }
void orc_stream_words(uint64_t seed, uint64_t stream, uint64_t first, int64_t n, uint32_t* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = oracle_philox_word(seed, stream, first + (uint64_t)i);
}

// ---- alias ---------------------------------------------------------------------------------------------------
void orc_alias_build(int sem, const double* dist, int64_t n, double power, double* prob, int64_t* alias) {
    std::vector<double> d(dist, dist + n);
    Alias t = sem == SEM_CPP ? alias_cpp(d) : alias_go(d, power);
    for (int64_t i = 0; i < n; ++i) { prob[i] = t.prob[i]; alias[i] = t.alias[i]; }
}

// ---- graph ---------------------------------------------------------------------------------------------------
void* orc_graph_create(int sem, int64_t V, int64_t E, const int64_t* row_off, const int32_t* col, const double* w,
                       int64_t max_line, int neg_method) {
    Graph* g = new Graph();
    g->sem = sem; g->V = V; g->E = E; g->max_line = max_line;
    g->off.assign(row_off, row_off + V + 1);
    g->col.assign(col, col + E);
    g->w.assign(w, w + E);
    g->build(neg_method);
    return g;
}
void orc_graph_free(void* h) { delete (Graph*)h; }
void orc_graph_set_field(void* h, const int32_t* field) { Graph* g = (Graph*)h; g->field.assign(field, field + g->V); }
void orc_graph_degrees(void* h, double* out_deg, double* in_deg) {
    Graph* g = (Graph*)h;
    memcpy(out_deg, g->out_deg.data(), g->V * sizeof(double));
    memcpy(in_deg, g->in_deg.data(), g->V * sizeof(double));
}
// which: 0 vertex_AT, 1 negative_AT, 2 context_AT (C++ only)
void orc_graph_alias(void* h, int which, double* prob, int64_t* alias) {
    Graph* g = (Graph*)h;
    Alias& t = which == 0 ? g->vertex_at : which == 1 ? g->negative_at : g->context_at;
    for (size_t i = 0; i < t.prob.size(); ++i) { prob[i] = t.prob[i]; alias[i] = t.alias[i]; }
}
void orc_sigmoid_table(void* h, double* out1001) { memcpy(out1001, ((Graph*)h)->sigmoid, sizeof(double) * 1001); }
double orc_fast_sigmoid(void* h, double x) { return ((Graph*)h)->fast_sigmoid(x); }

// which: 0 source, 1 negative, 2 target(arg[i]), 3 source+target pairs. Returns stream position afterwards.
uint64_t orc_sample(void* h, int which, uint64_t seed, uint64_t stream, int64_t n, const int64_t* arg, int64_t* out) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    for (int64_t i = 0; i < n; ++i) {
        if (which == 0) out[i] = g->source_sample(d);
        else if (which == 1) out[i] = g->negative_sample(d);
        else if (which == 2) out[i] = g->target_sample(arg[i], d);
        else { int64_t s = g->source_sample(d); out[2 * i] = s; out[2 * i + 1] = g->target_sample(s, d); }
    }
    return d.s.pos;
}

// mode 0: SkipGrams(window=w0); mode 1: ScaleSkipGrams(w0,w1). Returns #pairs.
int64_t orc_walk_pairs(void* h, uint64_t seed, uint64_t stream, int64_t start, int steps, int mode, int w0, int w1,
                       int64_t* walk, int64_t* walk_len, int64_t* pv, int64_t* pc, int64_t cap) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    std::vector<int64_t> wk, a, b;
    g->random_walk(start, steps, d, wk);
    *walk_len = (int64_t)wk.size();
    for (size_t i = 0; i < wk.size(); ++i) walk[i] = wk[i];
    if (mode == 0) g->skip_grams(wk, w0, d, a, b); else g->scale_skip_grams(wk, w0, w1, a, b);
    for (size_t i = 0; i < a.size() && (int64_t)i < cap; ++i) { pv[i] = a[i]; pc[i] = b[i]; }
    return (int64_t)a.size();
}

// ---- train loops (single stream == the reference at -threads 1) -----------------------------------------
// Every loop takes `total` directly (the CLI's sample_times*1e6 / sample_times*MaxLine / walk_times*V is the caller's
// business) and returns the number of stream words consumed.

// C++ LINE: src/model/LINE.cpp:100-195. order 1 passes the same table twice (Wc == Wv).
uint64_t orc_train_line_cpp(void* h, double* Wv, double* Wc, int dim, int K, double alpha, uint64_t total,
                            uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    CppSchedule sch(alpha, total);
    std::vector<double> be;
    unsigned long long jobs = total, count = 1;  // count starts at 1: jobs-1 samples (LINE.cpp:166,170)
    while (count < jobs) {
        int64_t v1 = g->source_sample(d);
        int64_t v2 = g->target_sample(v1, d);
        g->update_pair_cpp(Wv, Wc, v1, v2, dim, K, sch.cur, d, be);
        count++;
        sch.tick(count);
    }
    return d.s.pos;
}

// C++ BPR: src/model/BPR.cpp:55-107 (count starts at 0: `total` samples)
uint64_t orc_train_bpr_cpp(void* h, double* W, int dim, double alpha, uint64_t total, uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    CppSchedule sch(alpha, total);
    std::vector<double> a, b, c;
    unsigned long long count = 0;
    while (count < total) {
        int64_t v1 = g->source_sample(d);
        int64_t v2 = g->target_sample(v1, d);
        int64_t v3 = g->negative_sample(d);
        g->update_bpr_pair_cpp(W, v1, v2, v3, dim, sch.cur, d, a, b, c);
        count++;
        sch.tick(count);
    }
    return d.s.pos;
}

// C++ WARP: src/model/WARP.cpp:55-107. tries_out (optional) accumulates scanned negatives.
uint64_t orc_train_warp_cpp(void* h, double* W, int dim, double alpha, uint64_t total, uint64_t seed, uint64_t stream,
                            uint64_t* tries_out) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    CppSchedule sch(alpha, total);
    std::vector<double> a, b, c;
    unsigned long long count = 0, tries = 0;
    while (count < total) {
        int64_t v1 = g->source_sample(d);
        int64_t v2 = g->target_sample(v1, d);
        int64_t v3 = g->negative_sample(d);
        tries += g->update_warp_pair_cpp(W, v1, v2, v3, dim, sch.cur, d, a, b, c);
        count++;
        sch.tick(count);
    }
    if (tries_out) *tries_out = tries;
    return d.s.pos;
}

// C++ HOP-Rec: src/model/HBPR.cpp:63-130
uint64_t orc_train_hoprec_cpp(void* h, double* W, int dim, int walk_steps, double alpha, uint64_t total, uint64_t seed,
                              uint64_t stream) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    CppSchedule sch(alpha, total);
    std::vector<double> a, b, c;
    unsigned long long count = 0;
    while (count < total) {
        int64_t vid = g->source_sample(d);
        while (g->field[vid] != 0) vid = g->source_sample(d);
        int64_t cid1 = g->target_sample(vid, d);
        double margin = 1.0;
        for (int w = 1; w <= walk_steps; w++) {
            if (w != 1) {
                cid1 = g->target_sample(cid1, d);
                cid1 = g->target_sample(cid1, d);
            }
            int64_t nid = g->negative_sample(d);
            while (g->field[nid] != g->field[cid1]) nid = g->negative_sample(d);
            g->update_fbpr_pair_cpp(W, vid, cid1, nid, dim, sch.cur / w, margin / w, d, a, b, c);
        }
        count++;
        sch.tick(count);
    }
    return d.s.pos;
}

// C++ Skew-OPT: src/model/SkewOPT.cpp (SPR::Train; one table; negatives "no_degrees"; count from 0)
uint64_t orc_train_skewopt_cpp(void* h, double* W, int dim, double xi, double omega, int eta, double alpha, uint64_t total,
                               uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    CppSchedule sch(alpha, total);
    std::vector<double> a, b, c;
    unsigned long long count = 0;
    while (count < total) {
        int64_t v1 = g->source_sample(d);
        int64_t v2 = g->target_sample(v1, d);
        g->update_sbpr_pair_cpp(W, v1, v2, dim, xi, omega, eta, sch.cur, d, a, b, c);
        count++;
        sch.tick(count);
    }
    return d.s.pos;
}
uint64_t orc_update_sbpr_pair_cpp(void* h, double* W, int64_t v, int64_t ci, int dim, double xi, double omega, int eta,
                                  double alpha, uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    std::vector<double> a, b, c;
    g->update_sbpr_pair_cpp(W, v, ci, dim, xi, omega, eta, alpha, d, a, b, c);
    return d.s.pos;
}

// C++ MF: src/model/MF.cpp:50-98 (one table; negatives "no_degrees", MF.cpp:4-7; count from 0)
uint64_t orc_train_mf_cpp(void* h, double* W, int dim, int K, double reg, double alpha, uint64_t total, uint64_t seed,
                          uint64_t stream) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    CppSchedule sch(alpha, total);
    std::vector<double> be;
    unsigned long long count = 0;
    while (count < total) {
        int64_t v1 = g->source_sample(d);
        int64_t v2 = g->target_sample(v1, d);
        g->update_factorized_pair_cpp(W, W, v1, v2, dim, reg, K, sch.cur, d, be);
        count++;
        sch.tick(count);
    }
    return d.s.pos;
}
uint64_t orc_update_factorized_pair_cpp(void* h, double* W, int64_t v, int64_t c, int dim, double reg, int K, double alpha,
                                        uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    std::vector<double> be;
    g->update_factorized_pair_cpp(W, W, v, c, dim, reg, K, alpha, d, be);
    return d.s.pos;
}

// C++ HPE: src/model/HPE.cpp:93-147 -- UpdateCommunity(v1, v2) then UpdatePair with the roles swapped (v2, v1); count from 0
uint64_t orc_train_hpe_cpp(void* h, double* Wv, double* Wc, int dim, int walk_steps, int K, double reg, double alpha,
                           uint64_t total, uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    CppSchedule sch(alpha, total);
    std::vector<double> be;
    unsigned long long count = 0;
    while (count < total) {
        int64_t v1 = g->source_sample(d);
        int64_t v2 = g->target_sample(v1, d);
        g->update_community_cpp(Wv, Wc, v1, v2, dim, reg, walk_steps, K, sch.cur, d, be);
        g->update_pair_cpp(Wv, Wc, v2, v1, dim, K, sch.cur, d, be);
        count++;
        sch.tick(count);
    }
    return d.s.pos;
}
uint64_t orc_update_community_cpp(void* h, double* Wv, double* Wc, int64_t v, int64_t c, int dim, double reg, int walk_steps,
                                  int K, double alpha, uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    std::vector<double> be;
    g->update_community_cpp(Wv, Wc, v, c, dim, reg, walk_steps, K, alpha, d, be);
    return d.s.pos;
}

// C++ DeepWalk (walklets=0, src/model/DeepWalk.cpp:98-155) and Walklets (walklets=1, src/model/Walklets.cpp:6-64).
// The per-epoch Fisher-Yates shuffle pops libc rand() (here: the shuffle stream) in both; Walklets ignores the result.
uint64_t orc_train_walk_cpp(void* h, int walklets, double* Wv, double* Wc, int dim, int walk_times, int walk_steps,
                            int w0, int w1, int K, double alpha, uint64_t seed, uint64_t stream,
                            int64_t max_walks /* <0: all */, uint64_t* pairs_out) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    unsigned long long total = (unsigned long long)walk_times * g->V;
    double alpha_min = alpha * 0.0001, cur = alpha;
    unsigned long long count = 0, pairs = 0;
    std::vector<int64_t> keys(g->V), walk, pv, pc;
    std::vector<double> be;
    for (int t = 0; t < walk_times; ++t) {
        for (int64_t v = 0; v < g->V; ++v) keys[v] = v;
        for (int64_t v = 0; v < g->V; ++v) {
            int rdx = (int)(v + d.libc_rand() % (g->V - v));
            std::swap(keys[v], keys[rdx]);
        }
        for (int64_t v = 0; v < g->V; ++v) {
            if (max_walks >= 0 && (int64_t)count >= max_walks) goto done;
            g->random_walk(walklets ? v : keys[v], walk_steps, d, walk);
            if (walklets) g->scale_skip_grams(walk, w0, w1, pv, pc);
            else g->skip_grams(walk, w0, d, pv, pc);
            for (size_t i = 0; i < pv.size(); ++i) g->update_pair_cpp(Wv, Wc, pv[i], pc[i], dim, K, cur, d, be);
            pairs += pv.size();
            count++;
            if (count % MONITOR == 0) {
                cur = alpha * (1.0 - (double)(count) / total);
                if (cur < alpha_min) cur = alpha_min;
            }
        }
    }
done:
    if (pairs_out) *pairs_out = pairs;
    return d.s.pos;
}

// Go LINE: internal/models/line/line.go:73-150 (one worker). iterations = sampleTimes*MaxLine loop trips.
uint64_t orc_train_line_go(void* h, double* Wv, double* Wc, int dim, int order, int K, double alpha,
                           uint64_t iterations, uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    double alpha_min = alpha * 0.0001, cur = alpha;
    int64_t total = (int64_t)iterations, count = 0;
    std::vector<double> a, b, c;
    for (uint64_t it = 0; it < iterations; ++it) {
        int64_t s = g->source_sample(d);
        if (s == -1) continue;
        int64_t t = g->target_sample(s, d);
        if (t == -1) continue;
        if (order == 1) g->update_first_order_go(Wv, s, t, dim, K, cur, d, a, b);
        else g->update_pair_go(Wv, Wc, s, t, dim, K, cur, d, a, b, c);
        count++;
        if (count % MONITOR == 0) {
            cur = alpha * (1.0 - (double)count / (double)total);
            if (cur < alpha_min) cur = alpha_min;
        }
    }
    return d.s.pos;
}

// Go BPR: internal/models/bpr/bpr.go:61-131
uint64_t orc_train_bpr_go(void* h, double* Wv, double* Wc, int dim, double alpha, double lambda, uint64_t iterations,
                          uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    double alpha_min = alpha * 0.0001, cur = alpha;
    int64_t total = (int64_t)iterations, count = 0;
    for (uint64_t it = 0; it < iterations; ++it) {
        int64_t u = g->source_sample(d);
        if (u == -1) continue;
        int64_t p = g->target_sample(u, d);
        if (p == -1) continue;
        int64_t n = g->negative_sample(d);
        g->update_bpr_pair_go(Wv, Wc, u, p, n, dim, cur, lambda);
        count++;
        if (count % MONITOR == 0) {
            cur = alpha * (1.0 - (double)count / (double)total);
            if (cur < alpha_min) cur = alpha_min;
        }
    }
    return d.s.pos;
}

// Go CPR: internal/models/cpr/cpr.go:127-282 (one worker). `h` is the target-domain graph, `hs` the source-domain graph;
// U (user rows) and T (target-domain item rows) are trained, S (source-domain item rows) is only read. A user is looked up
// in BOTH graphs by its target-graph vid (cpr.go:53-59, :72-78, :148-169).
uint64_t orc_train_cpr_go(void* h, void* hs, double* U, double* T, const double* S, int dim, double alpha, double user_reg,
                          double item_reg, double margin, uint64_t iterations, uint64_t total_, uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h;
    Graph* gs = (Graph*)hs;
    Draws d(seed, stream);
    const double alpha_min = alpha * 0.0001;
    double cur = alpha;
    const int64_t total = (int64_t)total_;
    int64_t count = 0;
    std::vector<double> uv(dim), ug(dim), pg(dim), ng(dim);
    for (uint64_t it = 0; it < iterations; ++it) {
        const int64_t u = g->source_sample(d);
        const int64_t p = g->target_sample(u, d);
        if (p == -1) continue;  // (no count++: cpr.go:214-216)
        // transformUser (:127-172)
        for (int k = 0; k < dim; ++k) uv[k] = 0.0;
        double cnt = 0.0;
        for (int k = 0; k < dim; ++k) uv[k] += U[u * dim + k];
        cnt += 1.0;
        for (int64_t e = g->off[u]; e < g->off[u + 1]; ++e) {
            const double* r = T + (int64_t)g->col[e] * dim;
            for (int k = 0; k < dim; ++k) uv[k] += r[k];
            cnt += 1.0;
        }
        if (u < gs->V)
            for (int64_t e = gs->off[u]; e < gs->off[u + 1]; ++e) {
                const double* r = S + (int64_t)gs->col[e] * dim;
                for (int k = 0; k < dim; ++k) uv[k] += r[k];
                cnt += 1.0;
            }
        for (int k = 0; k < dim; ++k) uv[k] /= cnt;
        const int64_t n = g->negative_sample(d);
        double* ru = U + u * dim;
        double* rp = T + p * dim;
        double* rn = T + n * dim;
        double ps = 0.0, ns = 0.0;
        for (int k = 0; k < dim; ++k) {
            ps += uv[k] * rp[k];
            ns += uv[k] * rn[k];
        }
        const double diff = ps - ns;
        if (diff < margin) {
            const double gc = cur * g->fast_sigmoid(-(diff - margin));
            for (int k = 0; k < dim; ++k) {
                pg[k] = gc * uv[k];
                ng[k] = -gc * uv[k];
                ug[k] = gc * (rp[k] - rn[k]);
            }
            for (int k = 0; k < dim; ++k) {
                ru[k] -= cur * user_reg * ru[k];
                ru[k] += ug[k];
                rp[k] -= cur * item_reg * rp[k];
                rp[k] += pg[k];
                rn[k] -= cur * item_reg * rn[k];
                rn[k] += ng[k];
            }
        }
        count++;
        if (count % MONITOR == 0) {
            cur = alpha * (1.0 - (double)count / (double)total);
            if (cur < alpha_min) cur = alpha_min;
        }
    }
    return d.s.pos;
}

// Go TPR: internal/models/tpr/tpr.go:101-262 (one worker). `h` is the user-item graph, `hw` the item-word graph (an item
// is looked up there by its user-item vid, :108); U / I are the user / item rows (both V_ui), Wd the word rows (V_iw).
uint64_t orc_train_tpr_go(void* h, void* hw, double* U, double* I, double* Wd, int dim, double alpha, double lambda,
                          double text_weight, uint64_t iterations, uint64_t total_, uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h;
    Graph* gw = (Graph*)hw;
    Draws d(seed, stream);
    const double alpha_min = alpha * 0.0001;
    double cur = alpha;
    const int64_t total = (int64_t)total_;
    int64_t count = 0;
    std::vector<double> pv(dim), nv(dim), ug(dim), pg(dim), ng(dim);
    auto enriched = [&](int64_t item, std::vector<double>& out) {  // :101-121
        const double* ri = I + item * dim;
        for (int k = 0; k < dim; ++k) out[k] = (1.0 - text_weight) * ri[k];
        const int64_t nw = item < gw->V ? gw->off[item + 1] - gw->off[item] : 0;
        if (nw > 0) {
            for (int64_t e = gw->off[item]; e < gw->off[item + 1]; ++e) {
                const double* rw = Wd + (int64_t)gw->col[e] * dim;
                for (int k = 0; k < dim; ++k) out[k] += (text_weight / (double)nw) * rw[k];
            }
        } else {
            for (int k = 0; k < dim; ++k) out[k] = ri[k];
        }
    };
    auto push_words = [&](int64_t item, const std::vector<double>& grad) {  // :216-232
        const int64_t nw = item < gw->V ? gw->off[item + 1] - gw->off[item] : 0;
        if (nw <= 0) return;
        const double ww = text_weight / (double)nw;
        for (int64_t e = gw->off[item]; e < gw->off[item + 1]; ++e) {
            double* rw = Wd + (int64_t)gw->col[e] * dim;
            for (int k = 0; k < dim; ++k) rw[k] += ww * grad[k] - lambda * cur * rw[k];
        }
    };
    for (uint64_t it = 0; it < iterations; ++it) {
        const int64_t u = g->source_sample(d);
        if (u == -1) continue;
        const int64_t p = g->target_sample(u, d);
        if (p == -1) continue;
        const int64_t n = g->negative_sample(d);
        enriched(p, pv);
        enriched(n, nv);
        double* ru = U + u * dim;
        double ps = 0.0, ns = 0.0;
        for (int k = 0; k < dim; ++k) {
            ps += ru[k] * pv[k];
            ns += ru[k] * nv[k];
        }
        const double gc = cur * g->fast_sigmoid(ns - ps);
        for (int k = 0; k < dim; ++k) {
            ug[k] = gc * (pv[k] - nv[k]);
            pg[k] = gc * ru[k];
            ng[k] = -gc * ru[k];
        }
        for (int k = 0; k < dim; ++k) ru[k] += ug[k] - lambda * cur * ru[k];
        double* rp = I + p * dim;
        double* rn = I + n * dim;
        for (int k = 0; k < dim; ++k) {
            rp[k] += (1.0 - text_weight) * pg[k] - lambda * cur * rp[k];
            rn[k] += (1.0 - text_weight) * ng[k] - lambda * cur * rn[k];
        }
        push_words(p, pg);
        push_words(n, ng);
        count++;
        if (count % MONITOR == 0) {
            cur = alpha * (1.0 - (double)count / (double)total);
            if (cur < alpha_min) cur = alpha_min;
        }
    }
    return d.s.pos;
}

// Go DeepWalk: internal/models/deepwalk/deepwalk.go:61-141 (one worker). rand.Int63n(n) := shuffle-stream index draw.
uint64_t orc_train_deepwalk_go(void* h, double* Wv, double* Wc, int dim, int walk_times, int walk_steps, int window,
                               int K, double alpha, uint64_t seed, uint64_t stream, int64_t max_walks,
                               uint64_t* pairs_out) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    int64_t total = (int64_t)walk_times * g->V, count = 0;
    double alpha_min = alpha * 0.0001, cur = alpha;
    unsigned long long pairs = 0;
    std::vector<int64_t> keys(g->V), walk, pv, pc;
    std::vector<double> a, b, c;
    for (int t = 0; t < walk_times; ++t) {
        for (int64_t v = 0; v < g->V; ++v) keys[v] = v;
        for (int64_t v = 0; v < g->V; ++v) {
            int64_t j = v + d.shuffle_index(g->V - v);
            std::swap(keys[v], keys[j]);
        }
        for (int64_t v = 0; v < g->V; ++v) {
            if (max_walks >= 0 && count >= max_walks) goto done;
            g->random_walk(keys[v], walk_steps, d, walk);
            g->skip_grams(walk, window, d, pv, pc);
            for (size_t i = 0; i < pv.size(); ++i) g->update_pair_go(Wv, Wc, pv[i], pc[i], dim, K, cur, d, a, b, c);
            pairs += pv.size();
            count++;
            if (count % MONITOR == 0) {
                cur = alpha * (1.0 - (double)count / (double)total);
                if (cur < alpha_min) cur = alpha_min;
            }
        }
    }
done:
    if (pairs_out) *pairs_out = pairs;
    return d.s.pos;
}

// Diagnostic (tests/probes/go_walk_streams_probe.py): the DEVICE's work split of the Go DeepWalk loop, run sequentially -- W
// independent draw streams, walk i of an epoch on stream i % W, each stream with the device's tick rule for the shared
// schedule (device_core.cuh sched_tick) -- i.e. the Hogwild kernel minus the concurrency. Not a reference path.
uint64_t orc_train_deepwalk_go_streams(void* h, double* Wv, double* Wc, int dim, int walk_times, int walk_steps, int window,
                                       int K, double alpha, uint64_t seed, int W, uint64_t* pairs_out) {
    // W < 0: -W schedule slots fed by ONE draw stream (separates the effect of the schedule split from the draw split)
    const bool one_stream = W < 0;
    if (one_stream) W = -W;
    Graph* g = (Graph*)h;
    std::vector<Draws> d;
    for (int w = 0; w < W; ++w) d.emplace_back(seed, (uint64_t)w);
    const double total = (double)((int64_t)walk_times * g->V), alpha_min = alpha * 0.0001;
    std::vector<uint64_t> count((size_t)W, 0), next_tick((size_t)W, ((uint64_t)MONITOR + W - 1) / W);
    std::vector<double> cur((size_t)W, alpha);
    unsigned long long pairs = 0;
    std::vector<int64_t> keys(g->V), walk, pv, pc;
    std::vector<double> a, b, c;
    for (int t = 0; t < walk_times; ++t) {
        for (int64_t v = 0; v < g->V; ++v) keys[v] = v;
        for (int64_t v = 0; v < g->V; ++v) {
            int64_t j = v + d[0].shuffle_index(g->V - v);
            std::swap(keys[v], keys[j]);
        }
        for (int64_t v = 0; v < g->V; ++v) {
            const size_t w = (size_t)(v % W);
            Draws& dw = d[one_stream ? 0 : w];
            g->random_walk(keys[v], walk_steps, dw, walk);
            g->skip_grams(walk, window, dw, pv, pc);
            for (size_t i = 0; i < pv.size(); ++i) g->update_pair_go(Wv, Wc, pv[i], pc[i], dim, K, cur[w], dw, a, b, c);
            pairs += pv.size();
            count[w]++;
            if (count[w] >= next_tick[w]) {
                const uint64_t tk = count[w] * (uint64_t)W / MONITOR;
                cur[w] = std::max(alpha_min, alpha * (1.0 - (double)(tk * MONITOR) / total));
                next_tick[w] = ((tk + 1) * MONITOR + W - 1) / W;
            }
        }
    }
    if (pairs_out) *pairs_out = pairs;
    return 0;
}

// Go node2vec: internal/models/node2vec/node2vec.go:176-260 (one worker): DeepWalk.Train with the biased second-order walk.
uint64_t orc_train_node2vec_go(void* h, double* Wv, double* Wc, int dim, int walk_times, int walk_steps, int window,
                               int K, double alpha, double p, double q, uint64_t seed, uint64_t stream, int64_t max_walks,
                               uint64_t* pairs_out) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    int64_t total = (int64_t)walk_times * g->V, count = 0;
    double alpha_min = alpha * 0.0001, cur = alpha;
    unsigned long long pairs = 0;
    std::vector<int64_t> keys(g->V), walk, pv, pc;
    std::vector<double> a, b, c, bw;
    for (int t = 0; t < walk_times; ++t) {
        for (int64_t v = 0; v < g->V; ++v) keys[v] = v;
        for (int64_t v = 0; v < g->V; ++v) {
            int64_t j = v + d.shuffle_index(g->V - v);
            std::swap(keys[v], keys[j]);
        }
        for (int64_t v = 0; v < g->V; ++v) {
            if (max_walks >= 0 && count >= max_walks) goto done;
            g->biased_random_walk(keys[v], walk_steps, p, q, d, walk, bw);
            g->skip_grams(walk, window, d, pv, pc);
            for (size_t i = 0; i < pv.size(); ++i) g->update_pair_go(Wv, Wc, pv[i], pc[i], dim, K, cur, d, a, b, c);
            pairs += pv.size();
            count++;
            if (count % MONITOR == 0) {
                cur = alpha * (1.0 - (double)count / (double)total);
                if (cur < alpha_min) cur = alpha_min;
            }
        }
    }
done:
    if (pairs_out) *pairs_out = pairs;
    return d.s.pos;
}
// one biased walk from `start` (parity hook)
int64_t orc_biased_walk_go(void* h, int64_t start, int steps, double p, double q, uint64_t seed, uint64_t stream, int64_t* out,
                           uint64_t* words) {
    Graph* g = (Graph*)h;
    Draws d(seed, stream);
    std::vector<int64_t> walk;
    std::vector<double> bw;
    g->biased_random_walk(start, steps, p, q, d, walk, bw);
    for (size_t i = 0; i < walk.size(); ++i) out[i] = walk[i];
    if (words) *words = d.s.pos;
    return (int64_t)walk.size();
}

// ---- step-level entry points (unit parity with the compiled reference's Update*Pair) ------------------------
uint64_t orc_update_pair_cpp(void* h, double* Wv, double* Wc, int64_t v, int64_t c, int dim, int K, double alpha,
                             uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h; Draws d(seed, stream); std::vector<double> be;
    g->update_pair_cpp(Wv, Wc, v, c, dim, K, alpha, d, be);
    return d.s.pos;
}
uint64_t orc_update_bpr_pair_cpp(void* h, double* W, int64_t v, int64_t ci, int64_t cj, int dim, double alpha,
                                 uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h; Draws d(seed, stream); std::vector<double> a, b, c;
    g->update_bpr_pair_cpp(W, v, ci, cj, dim, alpha, d, a, b, c);
    return d.s.pos;
}
uint64_t orc_update_warp_pair_cpp(void* h, double* W, int64_t v, int64_t ci, int64_t cj, int dim, double alpha,
                                  uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h; Draws d(seed, stream); std::vector<double> a, b, c;
    g->update_warp_pair_cpp(W, v, ci, cj, dim, alpha, d, a, b, c);
    return d.s.pos;
}
uint64_t orc_update_fbpr_pair_cpp(void* h, double* W, int64_t v, int64_t ci, int64_t cj, int dim, double alpha,
                                  double margin, uint64_t seed, uint64_t stream) {
    Graph* g = (Graph*)h; Draws d(seed, stream); std::vector<double> a, b, c;
    g->update_fbpr_pair_cpp(W, v, ci, cj, dim, alpha, margin, d, a, b, c);
    return d.s.pos;
}

// Text edge list for the compiled reference's LoadEdgeList ("v<src> v<dst> <weight>" per line; bench/test helper).
int orc_write_edge_list(const char* path, const int64_t* src, const int64_t* dst, const double* w, int64_t n) {
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    std::vector<char> buf(1 << 22);
    setvbuf(f, buf.data(), _IOFBF, buf.size());
    for (int64_t i = 0; i < n; ++i) fprintf(f, "v%lld v%lld %g\n", (long long)src[i], (long long)dst[i], w[i]);
    fclose(f);
    return 0;
}

// ---- multi-threaded Hogwild leg (bench.py cpu_baseline "port"; the CPU side of the large-graph quality gate): LINE-2,
// C++ semantics, one stream per thread, and the reference's schedule exactly as LINE.cpp:162-191 runs it with `workers`
// threads: every worker counts from 1 to jobs = total / workers and, every MONITOR of ITS samples, reads the shared
// progress counter for its new alpha and then bumps it by MONITOR (the counter is a plain shared variable in the reference;
// relaxed atomics here).
double orc_time_line_cpp(void* h, double* Wv, double* Wc, int dim, int K, double alpha, uint64_t total, uint64_t seed,
                         int workers);

}  // extern "C"

#include <omp.h>

#include <atomic>
extern "C" double orc_time_line_cpp(void* h, double* Wv, double* Wc, int dim, int K, double alpha, uint64_t total,
                                    uint64_t seed, int workers) {
    Graph* g = (Graph*)h;
    unsigned long long jobs = total / workers;
    const double alpha_min = alpha * 0.0001;
    std::atomic<unsigned long long> current_sample{0};
    double t0 = omp_get_wtime();
#pragma omp parallel for num_threads(workers)
    for (int wk = 0; wk < workers; ++wk) {
        Draws d(seed, (uint64_t)wk);
        std::vector<double> be;
        unsigned long long count = 1;
        double cur = alpha;
        while (count < jobs) {
            int64_t v1 = g->source_sample(d);
            int64_t v2 = g->target_sample(v1, d);
            g->update_pair_cpp(Wv, Wc, v1, v2, dim, K, cur, d, be);
            count++;
            if (count % MONITOR == 0) {
                cur = alpha * (1.0 - (double)current_sample.load(std::memory_order_relaxed) / (double)total);
                current_sample.fetch_add(MONITOR, std::memory_order_relaxed);
                if (cur < alpha_min) cur = alpha_min;
            }
        }
    }
    return omp_get_wtime() - t0;
}
