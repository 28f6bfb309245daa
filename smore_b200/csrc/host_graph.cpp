#include "host_graph.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string_view>
#include <unordered_map>

namespace smore {

namespace {
constexpr double kPowerSample = 0.75;  // POWER_SAMPLE, src/proNet.h:33
}

AliasHost alias_method_cpp(const double* dist, int64_t n) {
    AliasHost t;
    t.prob.assign((size_t)n, 0.0);
    t.alias.assign((size_t)n, -1);
    // normalisation exactly as written in the reference: sum of pow, norm = n / sum, pow * norm
    double sum = 0;
    for (int64_t i = 0; i < n; ++i) sum += std::pow(dist[i], kPowerSample);
    const double norm = (double)(size_t)n / sum;
    std::vector<double> np((size_t)n);
    for (int64_t i = 0; i < n; ++i) np[(size_t)i] = std::pow(dist[i], kPowerSample) * norm;
    // two LIFO stacks filled in index order; the pop/pair/push order determines the table
    std::vector<int64_t> small_block, large_block;
    small_block.reserve((size_t)n);
    large_block.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        if (np[(size_t)i] < 1) small_block.push_back(i);
        else large_block.push_back(i);
    }
    while (!small_block.empty() && !large_block.empty()) {
        const int64_t s = small_block.back();
        small_block.pop_back();
        const int64_t l = large_block.back();
        large_block.pop_back();
        t.alias[(size_t)s] = l;
        t.prob[(size_t)s] = np[(size_t)s];
        np[(size_t)l] = np[(size_t)l] + np[(size_t)s] - 1;
        if (np[(size_t)l] < 1) small_block.push_back(l);
        else large_block.push_back(l);
    }
    for (int64_t l : large_block) t.prob[(size_t)l] = 1.0;
    for (int64_t s : small_block) t.prob[(size_t)s] = 1.0;
    return t;
}

AliasHost alias_method_go(const double* dist, int64_t n, double power) {
    AliasHost t;
    t.prob.assign((size_t)n, 0.0);
    t.alias.assign((size_t)n, 0);
    if (n == 0) return t;
    std::vector<double> norm((size_t)n);
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        norm[(size_t)i] = dist[i] > 0 ? std::pow(dist[i], power) : 0.0;
        sum += norm[(size_t)i];
    }
    if (sum == 0) {
        for (int64_t i = 0; i < n; ++i) {
            t.prob[(size_t)i] = 1.0;
            t.alias[(size_t)i] = i;
        }
        return t;
    }
    for (int64_t i = 0; i < n; ++i) norm[(size_t)i] = norm[(size_t)i] * (double)n / sum;
    std::vector<int64_t> small, large;
    small.reserve((size_t)n);
    large.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        if (norm[(size_t)i] < 1.0) small.push_back(i);
        else large.push_back(i);
    }
    while (!small.empty() && !large.empty()) {
        const int64_t l = small.back();
        small.pop_back();
        const int64_t g = large.back();
        large.pop_back();
        t.prob[(size_t)l] = norm[(size_t)l];
        t.alias[(size_t)l] = g;
        norm[(size_t)g] = norm[(size_t)g] + norm[(size_t)l] - 1.0;
        if (norm[(size_t)g] < 1.0) small.push_back(g);
        else large.push_back(g);
    }
    for (int64_t g : large) { t.prob[(size_t)g] = 1.0; t.alias[(size_t)g] = g; }
    for (int64_t l : small) { t.prob[(size_t)l] = 1.0; t.alias[(size_t)l] = l; }
    return t;
}

PackedAlias pack_alias(double prob, int64_t alias, uint32_t self_id) {
    PackedAlias p;
    const double scaled = std::ceil(prob * 4294967296.0);  // exact: power-of-two scaling, then ceil
    if (!(scaled < 4294967296.0) || alias < 0) {
        // u = k*2^-32 < prob holds for every 32-bit k: the alias branch is unreachable
        p.thr = 0xFFFFFFFFu;
        p.alias = self_id;
    } else {
        p.thr = scaled <= 0.0 ? 0u : (uint32_t)scaled;
        p.alias = (uint32_t)alias;
    }
    return p;
}

bool load_edge_list(const char* path, bool undirected, EdgeList& out, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        err = std::string("cannot open edge list: ") + path;
        return false;
    }
    std::fseek(f, 0, SEEK_END);
    long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::string buf;
    buf.resize((size_t)size);
    if (size > 0 && std::fread(&buf[0], 1, (size_t)size, f) != (size_t)size) {
        std::fclose(f);
        err = std::string("short read: ") + path;
        return false;
    }
    std::fclose(f);

    std::unordered_map<std::string_view, int32_t> ids;
    std::vector<std::string_view> name_views;
    std::vector<int32_t> src, dst;
    std::vector<double> wt;
    auto intern = [&](std::string_view s) -> int32_t {
        auto it = ids.find(s);
        if (it != ids.end()) return it->second;
        int32_t id = (int32_t)name_views.size();
        ids.emplace(s, id);
        name_views.push_back(s);
        return id;
    };
    const char* p = buf.data();
    const char* end = p + buf.size();
    auto is_space = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; };
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        std::string_view tok[3];
        int nt = 0;
        const char* q = p;
        while (q < eol && nt < 3) {
            while (q < eol && is_space(*q)) ++q;
            if (q >= eol) break;
            const char* s = q;
            while (q < eol && !is_space(*q)) ++q;
            tok[nt++] = std::string_view(s, (size_t)(q - s));
        }
        if (nt == 3) {
            std::string num(tok[2]);
            char* endp = nullptr;
            double w = std::strtod(num.c_str(), &endp);
            if (endp != num.c_str() && *endp == '\0') {
                int32_t a = intern(tok[0]);  // ids in first-appearance order, source before target
                int32_t b = intern(tok[1]);
                src.push_back(a);
                dst.push_back(b);
                wt.push_back(w);
            }
        }
        p = eol + 1;
    }
    const int64_t V = (int64_t)name_views.size();
    const int64_t L = (int64_t)src.size();
    out.n_lines = L;
    out.names.clear();
    out.names.reserve((size_t)V);
    for (auto& sv : name_views) out.names.emplace_back(sv);
    // CSR in insertion order: the forward entry, then (undirected) the reverse entry, line by line
    out.row_off.assign((size_t)V + 1, 0);
    for (int64_t i = 0; i < L; ++i) {
        out.row_off[(size_t)src[(size_t)i] + 1]++;
        if (undirected) out.row_off[(size_t)dst[(size_t)i] + 1]++;
    }
    for (int64_t v = 0; v < V; ++v) out.row_off[(size_t)v + 1] += out.row_off[(size_t)v];
    const int64_t E = out.row_off[(size_t)V];
    out.col.assign((size_t)E, 0);
    out.w.assign((size_t)E, 0.0);
    std::vector<int64_t> cursor(out.row_off.begin(), out.row_off.end() - 1);
    for (int64_t i = 0; i < L; ++i) {
        int64_t e = cursor[(size_t)src[(size_t)i]]++;
        out.col[(size_t)e] = dst[(size_t)i];
        out.w[(size_t)e] = wt[(size_t)i];
        if (undirected) {
            e = cursor[(size_t)dst[(size_t)i]]++;
            out.col[(size_t)e] = src[(size_t)i];
            out.w[(size_t)e] = wt[(size_t)i];
        }
    }
    return true;
}

}  // namespace smore
