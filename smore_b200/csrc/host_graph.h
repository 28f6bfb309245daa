// Host-side graph construction: CSR + the reference's alias tables, built bit-exactly, then packed for the device.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace smore {

struct AliasHost {
    std::vector<double> prob;
    std::vector<int64_t> alias;
};

// proNet::AliasMethod (src/proNet.cpp:544-620): Vose with two LIFO stacks; the power argument is ignored by the
// reference and 0.75 always applied; alias stays -1 on prob == 1 entries.
AliasHost alias_method_cpp(const double* dist, int64_t n);
// BuildAliasMethod (pkg/pronet/alias.go:10-90): power honoured, zero weights stay 0, all-zero -> uniform, alias = self.
AliasHost alias_method_go(const double* dist, int64_t n, double power);

// Same algorithm writing into caller-provided slices, with caller-provided scratch (the per-vertex context tables of
// proNet::BuildAliasMethod, src/proNet.cpp:519-536, are built by many host threads at once).
struct AliasScratch {
    std::vector<double> np;
    std::vector<int64_t> small_block, large_block;
};
void alias_method_cpp_into(const double* dist, int64_t n, double* prob, int64_t* alias, AliasScratch& scratch);
// pow(x, 0.75) through a table for small integer weights (the same libm call, made once per distinct value).
double pow075(double x);

// Host worker threads (SMORE_HOST_THREADS, default: hardware concurrency, at most 64). fn(begin, end, thread) is called
// on contiguous slices of [0, n); `grain` = smallest slice worth a thread.
int host_threads();
template <class F>
void parallel_for(int64_t n, int64_t grain, F&& fn);

// Text formatting of embedding rows exactly as the reference writers do it: `name v0 v1 ...\n`, format 0 = ostream
// default ("%g", src/model/LINE.cpp:35-38), 1 = Go "%.6f" (internal/models/line/line.go:226). Appends to `out`.
void format_rows(const double* rows, int64_t n, int dim, const std::string* names /* null: decimal ids */,
                 int64_t first_id, int format, std::string& out);

struct PackedAlias {
    uint32_t thr;
    uint32_t alias;
};
// thr = ceil(prob * 2^32); entries that can never take the alias (prob > 1 - 2^-32) get alias = self_id.
PackedAlias pack_alias(double prob, int64_t alias, uint32_t self_id);

struct EdgeList {
    std::vector<int64_t> row_off;
    std::vector<int32_t> col;
    std::vector<double> w;
    std::vector<std::string> names;
    int64_t n_lines = 0;  // valid edge lines read
};
// Reference-compatible text ingest (src/proNet.cpp:158-224; pronet.go:128-165). Returns false + message on IO error.
bool load_edge_list(const char* path, bool undirected, EdgeList& out, std::string& err);

}  // namespace smore

// ---- parallel_for (header-only: it is a template) ----------------------------------------------------------------
#include <algorithm>
#include <thread>
namespace smore {
template <class F>
void parallel_for(int64_t n, int64_t grain, F&& fn) {
    if (n <= 0) return;
    int64_t t = std::min<int64_t>(host_threads(), (n + grain - 1) / std::max<int64_t>(grain, 1));
    if (t <= 1) {
        fn((int64_t)0, n, 0);
        return;
    }
    std::vector<std::thread> th;
    th.reserve((size_t)t);
    const int64_t per = (n + t - 1) / t;
    for (int64_t k = 0; k < t; ++k) {
        const int64_t b = k * per, e = std::min(n, b + per);
        if (b >= e) break;
        th.emplace_back([&fn, b, e, k]() { fn(b, e, (int)k); });
    }
    for (auto& x : th) x.join();
}
}  // namespace smore
