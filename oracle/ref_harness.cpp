// TEST INFRASTRUCTURE ONLY -- C-ABI harness around the UNMODIFIED compiled reference (C++ tree).
//
// Built by oracle/Makefile from the reference sources where they lie under /root/reference/src
// (proNet.cpp, util.cpp, model/{LINE,DeepWalk,Walklets,BPR,WARP,HBPR,HPE,MF,SkewOPT}.cpp) with random.cpp replaced
// by ref_shim.cpp; output goes to oracle/_ref/ only. Nothing here is copied from the reference: this
// file only CALLS its public classes (src/proNet.h:109-269, src/model/*.h).
//
// Used by tests/ (to pin oracle/smore_oracle.cpp and to generate tests/golden/*) and by
// bench.py --impl reference / cpu_baseline (timing leg).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "model/BPR.h"
#include "model/DeepWalk.h"
#include "model/HBPR.h"
#include "model/HPE.h"
#include "model/LINE.h"
#include "model/MF.h"
#include "model/SkewOPT.h"
#include "model/WARP.h"
#include "model/Walklets.h"

extern "C" void ref_shim_seed(uint64_t seed, uint64_t stream_base);
extern "C" uint64_t ref_shim_pos(void);

namespace {

enum Kind { K_LINE = 0, K_DEEPWALK = 1, K_WALKLETS = 2, K_BPR = 3, K_WARP = 4, K_HOPREC = 5, K_HPE = 6, K_MF = 7, K_SKEWOPT = 8 };

struct Ref {
    int kind;
    LINE* line = nullptr;
    DeepWalk* dw = nullptr;  // also Walklets (subclass)
    Walklets* wl = nullptr;
    BPR* bpr = nullptr;
    WARP* warp = nullptr;
    HBPR* hbpr = nullptr;
    HPE* hpe = nullptr;
    MF* mf = nullptr;
    SPR* spr = nullptr;
    int dim = 0;
    int order = 2;

    proNet& pnet() {
        switch (kind) {
            case K_LINE: return line->pnet;
            case K_DEEPWALK: return dw->pnet;
            case K_WALKLETS: return wl->pnet;
            case K_BPR: return bpr->pnet;
            case K_WARP: return warp->pnet;
            case K_HPE: return hpe->pnet;
            case K_MF: return mf->pnet;
            case K_SKEWOPT: return spr->pnet;
            default: return hbpr->pnet;
        }
    }
    // table 0 = vertex, 1 = context
    std::vector<std::vector<double>>* table(int t) {
        switch (kind) {
            case K_LINE:
                if (order == 1) return t == 0 ? &line->w_vertex_o1 : nullptr;
                return t == 0 ? &line->w_vertex : &line->w_context;
            case K_DEEPWALK: return t == 0 ? &dw->w_vertex : &dw->w_context;
            case K_WALKLETS: return t == 0 ? &wl->w_vertex : &wl->w_context;
            case K_BPR: return t == 0 ? &bpr->w_vertex : nullptr;
            case K_WARP: return t == 0 ? &warp->w_vertex : nullptr;
            case K_HPE: return t == 0 ? &hpe->w_vertex : &hpe->w_context;
            case K_MF: return t == 0 ? &mf->w_vertex : nullptr;
            case K_SKEWOPT: return t == 0 ? &spr->w_vertex : nullptr;
            default: return t == 0 ? &hbpr->w_vertex : &hbpr->w_context;
        }
    }
};

}  // namespace

extern "C" {

void* ref_new(int kind) {
    Ref* r = new Ref();
    r->kind = kind;
    switch (kind) {
        case K_LINE: r->line = new LINE(); break;
        case K_DEEPWALK: r->dw = new DeepWalk(); break;
        case K_WALKLETS: r->wl = new Walklets(); break;
        case K_BPR: r->bpr = new BPR(); break;
        case K_WARP: r->warp = new WARP(); break;
        case K_HOPREC: r->hbpr = new HBPR(); break;
        case K_HPE: r->hpe = new HPE(); break;
        case K_MF: r->mf = new MF(); break;
        case K_SKEWOPT: r->spr = new SPR(); break;
        default: delete r; return nullptr;
    }
    return r;
}

void ref_free(void* h) {
    Ref* r = (Ref*)h;
    delete r->line; delete r->dw; delete r->wl; delete r->bpr; delete r->warp; delete r->hbpr; delete r->hpe; delete r->mf; delete r->spr;
    delete r;
}

void ref_seed(uint64_t seed, uint64_t stream_base) { ref_shim_seed(seed, stream_base); }
uint64_t ref_stream_pos(void) { return ref_shim_pos(); }

void ref_load_edge_list(void* h, const char* path, int undirected) {
    Ref* r = (Ref*)h;
    r->pnet().LoadEdgeList(std::string(path), undirected != 0);
}

void ref_load_field(void* h, const char* path) {
    Ref* r = (Ref*)h;
    r->pnet().LoadFieldMeta(std::string(path));
}

void ref_init(void* h, int dim, int order) {
    Ref* r = (Ref*)h;
    r->dim = dim;
    r->order = order == 1 ? 1 : 2;
    switch (r->kind) {
        case K_LINE: r->line->Init(dim, order); break;
        case K_DEEPWALK: r->dw->Init(dim); break;
        case K_WALKLETS: r->wl->Init(dim); break;
        case K_BPR: r->bpr->Init(dim); break;
        case K_WARP: r->warp->Init(dim); break;
        case K_HPE: r->hpe->Init(dim); break;
        case K_MF: r->mf->Init(dim); break;
        case K_SKEWOPT: r->spr->Init(dim); break;
        default: r->hbpr->Init(dim); break;
    }
}

int64_t ref_num_vertices(void* h) { return ((Ref*)h)->pnet().MAX_vid; }
int64_t ref_num_lines(void* h) { return (int64_t)((Ref*)h)->pnet().MAX_line; }

// CSR + names as the reference built them (ingest parity)
void ref_get_csr(void* h, int64_t* row_off, int32_t* col, double* w) {
    proNet& p = ((Ref*)h)->pnet();
    for (long v = 0; v < p.MAX_vid; ++v) row_off[v] = p.vertex[v].offset;
    row_off[p.MAX_vid] = (int64_t)p.context.size();
    for (size_t e = 0; e < p.context.size(); ++e) {
        col[e] = (int32_t)p.context[e].vid;
        w[e] = p.context[e].in_degree;
    }
}

void ref_get_degrees(void* h, double* out_deg, double* in_deg) {
    proNet& p = ((Ref*)h)->pnet();
    for (long v = 0; v < p.MAX_vid; ++v) { out_deg[v] = p.vertex[v].out_degree; in_deg[v] = p.vertex[v].in_degree; }
}

const char* ref_vertex_name(void* h, int64_t vid) { return ((Ref*)h)->pnet().vertex_hash.keys[vid]; }

int32_t ref_field_of(void* h, int64_t vid) { return ((Ref*)h)->pnet().field[vid].fields[0]; }

// which: 0 vertex_AT, 1 negative_AT, 2 context_AT
int64_t ref_alias_size(void* h, int which) {
    proNet& p = ((Ref*)h)->pnet();
    return which == 0 ? p.vertex_AT.size() : which == 1 ? p.negative_AT.size() : p.context_AT.size();
}
void ref_get_alias(void* h, int which, double* prob, int64_t* alias) {
    proNet& p = ((Ref*)h)->pnet();
    std::vector<AliasTable>& t = which == 0 ? p.vertex_AT : which == 1 ? p.negative_AT : p.context_AT;
    for (size_t i = 0; i < t.size(); ++i) { prob[i] = t[i].prob; alias[i] = t[i].alias; }
}

// Stand-alone AliasMethod on a caller distribution (src/proNet.cpp:544-620).
void ref_alias_method(const double* dist, int64_t n, double power, double* prob, int64_t* alias) {
    static proNet* p = new proNet();
    std::vector<double> d(dist, dist + n);
    std::vector<AliasTable> t = p->AliasMethod(d, power);
    for (int64_t i = 0; i < n; ++i) { prob[i] = t[i].prob; alias[i] = t[i].alias; }
}

double ref_fast_sigmoid(void* h, double x) { return ((Ref*)h)->pnet().fastSigmoid(x); }
void ref_sigmoid_table(void* h, double* out1001) {
    proNet& p = ((Ref*)h)->pnet();
    // entry 1000 is the reference's one-past-the-end write (src/proNet.cpp:54-57); read it the way fastSigmoid(8.0) does.
    for (int i = 0; i < 1000; ++i) out1001[i] = p.cached_sigmoid[i];
    out1001[1000] = p.fastSigmoid(8.0);
}

// Sampler replay. which: 0 SourceSample, 1 NegativeSample, 2 TargetSample(arg[i]), 3 Source+Target pairs (out = 2n)
void ref_sample(void* h, int which, int64_t n, const int64_t* arg, int64_t* out) {
    proNet& p = ((Ref*)h)->pnet();
    for (int64_t i = 0; i < n; ++i) {
        if (which == 0) out[i] = p.SourceSample();
        else if (which == 1) out[i] = p.NegativeSample();
        else if (which == 2) out[i] = p.TargetSample((long)arg[i]);
        else { long s = p.SourceSample(); out[2 * i] = s; out[2 * i + 1] = p.TargetSample(s); }
    }
}

// RandomWalk + SkipGrams / ScaleSkipGrams replay: returns #pairs, fills walk (<= steps+1) and pairs.
int64_t ref_walk_pairs(void* h, int64_t start, int steps, int mode, int w0, int w1, int64_t* walk, int64_t* walk_len,
                       int64_t* pv, int64_t* pc, int64_t cap) {
    proNet& p = ((Ref*)h)->pnet();
    std::vector<long> wk = p.RandomWalk((long)start, steps);
    *walk_len = (int64_t)wk.size();
    for (size_t i = 0; i < wk.size(); ++i) walk[i] = wk[i];
    std::vector<std::vector<long>> tr = mode == 0 ? p.SkipGrams(wk, w0, 0) : p.ScaleSkipGrams(wk, w0, w1, 0);
    int64_t n = (int64_t)tr[0].size();
    for (int64_t i = 0; i < n && i < cap; ++i) { pv[i] = tr[0][i]; pc[i] = tr[1][i]; }
    return n;
}

void ref_set_rows(void* h, int table, const double* src) {
    Ref* r = (Ref*)h;
    auto* t = r->table(table);
    if (!t) return;
    for (size_t v = 0; v < t->size(); ++v)
        for (int d = 0; d < r->dim; ++d) (*t)[v][d] = src[v * r->dim + d];
}

void ref_get_rows(void* h, int table, double* dst) {
    Ref* r = (Ref*)h;
    auto* t = r->table(table);
    if (!t) return;
    for (size_t v = 0; v < t->size(); ++v)
        for (int d = 0; d < r->dim; ++d) dst[v * r->dim + d] = (*t)[v][d];
}

// Step-level entry points (the proNet core, not the model loop) on the model's own tables.
void ref_update_pair(void* h, int64_t v, int64_t c, int K, double alpha) {
    Ref* r = (Ref*)h;
    auto* tv = r->table(0);
    auto* tc = r->table(1) ? r->table(1) : tv;
    r->pnet().UpdatePair(*tv, *tc, (long)v, (long)c, r->dim, K, alpha);
}
void ref_update_bpr_pair(void* h, int64_t v, int64_t ci, int64_t cj, double alpha) {
    Ref* r = (Ref*)h;
    auto* tv = r->table(0);
    r->pnet().UpdateBPRPair(*tv, *tv, (long)v, (long)ci, (long)cj, r->dim, 0.0, alpha);
}
void ref_update_warp_pair(void* h, int64_t v, int64_t ci, int64_t cj, double alpha) {
    Ref* r = (Ref*)h;
    auto* tv = r->table(0);
    r->pnet().UpdateWARPPair(*tv, *tv, (long)v, (long)ci, (long)cj, r->dim, alpha);
}
void ref_update_fbpr_pair(void* h, int64_t v, int64_t ci, int64_t cj, double alpha, double margin) {
    Ref* r = (Ref*)h;
    auto* tv = r->table(0);
    r->pnet().UpdateFBPRPair(*tv, *tv, (long)v, (long)ci, (long)cj, r->dim, alpha, margin);
}

// Model Train() exactly as the reference CLIs call it (cli/line.cpp:76, cli/deepwalk.cpp, cli/walklets.cpp,
// cli/bpr.cpp, cli/warp.cpp, cli/hoprec.cpp). a/b/c/d are the model-specific integer args.
void ref_train(void* h, int a, int b, int c, int d, int e, double alpha, int workers) {
    Ref* r = (Ref*)h;
    switch (r->kind) {
        case K_LINE: r->line->Train(a /*sample_times*/, b /*K*/, alpha, workers); break;
        case K_DEEPWALK: r->dw->Train(a /*walk_times*/, b /*walk_steps*/, c /*window*/, d /*K*/, alpha, workers); break;
        case K_WALKLETS: r->wl->Train(a, b, c /*wmin*/, d /*wmax*/, e /*K*/, alpha, workers); break;
        case K_BPR: r->bpr->Train(a /*sample_times*/, b /*K (unused)*/, alpha, 0.01, workers); break;
        case K_WARP: r->warp->Train(a, b, alpha, 0.01, workers); break;
        default: r->hbpr->Train(a /*sample_times*/, b /*walk_steps*/, alpha, workers); break;
    }
}

// HPE::Train (src/model/HPE.cpp:93-147) as cli/hpe.cpp:76 calls it
void ref_train_hpe(void* h, int sample_times, int walk_steps, int K, double reg, double alpha, int workers) {
    Ref* r = (Ref*)h;
    if (r->kind == K_HPE) r->hpe->Train(sample_times, walk_steps, K, reg, alpha, workers);
}

// SPR::Train (src/model/SkewOPT.cpp) as cli/skewopt.cpp calls it
void ref_train_skewopt(void* h, int sample_times, int K, double alpha, double reg, double xi, double omega, int eta, int workers) {
    Ref* r = (Ref*)h;
    if (r->kind == K_SKEWOPT) r->spr->Train(sample_times, K, alpha, reg, xi, omega, eta, workers);
}

// proNet::UpdateSBPRPair (src/proNet.cpp:1517-1566) with one table in both roles, as SPR::Train passes it
void ref_update_sbpr_pair(void* h, int64_t v, int64_t ci, double xi, double omega, int eta, double alpha) {
    Ref* r = (Ref*)h;
    r->pnet().UpdateSBPRPair(*r->table(0), *r->table(0), (long)v, (long)ci, r->dim, 0.01, xi, omega, eta, alpha);
}

// MF::Train (src/model/MF.cpp:50-98) as cli/mf.cpp:65 calls it
void ref_train_mf(void* h, int sample_times, int K, double alpha, double reg, int workers) {
    Ref* r = (Ref*)h;
    if (r->kind == K_MF) r->mf->Train(sample_times, K, alpha, reg, workers);
}

// proNet::UpdateFactorizedPair (src/proNet.cpp:2591-2614) with one table in both roles, as MF::Train passes it
void ref_update_factorized_pair(void* h, int64_t v, int64_t c, double reg, int K, double alpha) {
    Ref* r = (Ref*)h;
    r->pnet().UpdateFactorizedPair(*r->table(0), *r->table(0), (long)v, (long)c, r->dim, reg, K, alpha);
}

// proNet::UpdateCommunity (src/proNet.cpp:3018-3054) on the model's own tables
void ref_update_community(void* h, int64_t v, int64_t c, double reg, int walk_steps, int K, double alpha) {
    Ref* r = (Ref*)h;
    r->pnet().UpdateCommunity(*r->table(0), *r->table(1), (long)v, (long)c, r->dim, reg, walk_steps, K, alpha);
}

void ref_save_weights(void* h, const char* path) {
    Ref* r = (Ref*)h;
    switch (r->kind) {
        case K_LINE: r->line->SaveWeights(path); break;
        case K_DEEPWALK: r->dw->SaveWeights(path); break;
        case K_WALKLETS: r->wl->SaveWeights(path); break;
        case K_BPR: r->bpr->SaveWeights(path); break;
        case K_WARP: r->warp->SaveWeights(path); break;
        case K_HPE: r->hpe->SaveWeights(path); break;
        case K_MF: r->mf->SaveWeights(path); break;
        case K_SKEWOPT: r->spr->SaveWeights(path); break;
        default: r->hbpr->SaveWeights(path); break;
    }
}

}  // extern "C"
