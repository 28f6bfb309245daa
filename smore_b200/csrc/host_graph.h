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
