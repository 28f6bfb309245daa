// C-ABI implementation (include/smore_b200.h): handles, HBM layout, host-side graph build. No CPU fallback anywhere.
// The trainers (kernel launches) are instantiated per element type in train_*.cu.
#include "host_common.h"

#include <charconv>
#include <string_view>
#include <unordered_map>

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
int g_device = -1;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

int ensure_device() {
    static thread_local int checked_dev = -2;
    if (checked_dev == g_device && g_device >= 0) return SMORE_OK;  // cudaGetDeviceProperties costs milliseconds
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(SMORE_E_CUDA, "no CUDA device available (%s); smore_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    int dev = g_device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(SMORE_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
    CU(cudaSetDevice(dev));
    g_device = dev;
    checked_dev = dev;
    return SMORE_OK;
}


namespace {

int upload_graph(smore_graph_s* g) {
    const int64_t V = g->V, E = g->E;
    if (int rc = dev_alloc_copy(&g->d_row_off, g->row_off.data(), (size_t)V + 1)) return rc;
    if (int rc = dev_alloc_copy(&g->d_col, g->col.data(), (size_t)E)) return rc;
    auto pack_table = [&](const AliasHost& t, std::vector<uint2>& out) {
        out.resize(t.prob.size());
        parallel_for((int64_t)t.prob.size(), 1 << 16, [&](int64_t b, int64_t e, int) {
            for (int64_t i = b; i < e; ++i) {
                PackedAlias p = pack_alias(t.prob[(size_t)i], t.alias[(size_t)i], (uint32_t)i);
                out[(size_t)i] = make_uint2(p.thr, p.alias);
            }
        });
    };
    std::vector<uint2> packed;
    pack_table(g->vertex_at, packed);
    if (int rc = dev_alloc_copy(&g->d_vat, packed.data(), packed.size())) return rc;
    pack_table(g->negative_at, packed);
    if (int rc = dev_alloc_copy(&g->d_nat, packed.data(), packed.size())) return rc;
    if (g->sem == SMORE_SEM_CPP) {
        // context table: alias entries are already vertex ids (src/proNet.cpp:530-534); "self" = the entry's own target
        packed.resize((size_t)E);
        parallel_for(E, 1 << 16, [&](int64_t b, int64_t en, int) {
            for (int64_t e = b; e < en; ++e) {
                PackedAlias p = pack_alias(g->ctx_at.prob[(size_t)e], g->ctx_at.alias[(size_t)e], (uint32_t)g->col[(size_t)e]);
                packed[(size_t)e] = make_uint2(p.thr, p.alias);
            }
        });
        if (int rc = dev_alloc_copy(&g->d_cat, packed.data(), packed.size())) return rc;
    } else {
        std::vector<double> prefix((size_t)E);
        parallel_for(V, 1 << 14, [&](int64_t vb, int64_t ve, int) {
            for (int64_t v = vb; v < ve; ++v) {
                double cum = 0.0;  // pronet.go:274-281: cumWeight += w in adjacency order
                for (int64_t e = g->row_off[(size_t)v]; e < g->row_off[(size_t)v + 1]; ++e) {
                    cum += g->w[(size_t)e];
                    prefix[(size_t)e] = cum;
                }
            }
        });
        if (int rc = dev_alloc_copy(&g->d_prefix, prefix.data(), prefix.size())) return rc;
    }
    g->field.assign((size_t)V, 0);
    if (int rc = dev_alloc_copy(&g->d_field, g->field.data(), (size_t)V)) return rc;
    // sigmoid LUT: InitSigmoid (src/proNet.cpp:52-60) incl. its one-past-the-end entry; pronet.go:90-95
    double lut[kSigmoidTable + 1];
    float lutf[kSigmoidTable + 1];
    for (int i = 0; i != kSigmoidTable + 1; i++) {
        double x = i * 2.0 * 8.0 / kSigmoidTable - 8.0;
        lut[i] = 1.0 / (1.0 + std::exp(-x));
        lutf[i] = (float)lut[i];
    }
    if (int rc = dev_alloc_copy(&g->d_lut64, lut, kSigmoidTable + 1)) return rc;
    if (int rc = dev_alloc_copy(&g->d_lut32, lutf, kSigmoidTable + 1)) return rc;
    return SMORE_OK;
}

int build_graph(smore_graph_s* g) {
    const int64_t V = g->V, E = g->E;
    if (V <= 0 || V >= (1ll << 31)) return fail(SMORE_E_INVALID, "V=%lld out of range", (long long)V);
    for (int64_t e = 0; e < E; ++e)
        if (g->col[(size_t)e] < 0 || g->col[(size_t)e] >= V) return fail(SMORE_E_INVALID, "col[%lld] out of range", (long long)e);
    for (int64_t v = 0; v < V; ++v)
        if (g->row_off[(size_t)v + 1] < g->row_off[(size_t)v] || g->row_off[(size_t)v + 1] - g->row_off[(size_t)v] >= (1ll << 32))
            return fail(SMORE_E_INVALID, "row_off not monotone / degree too large at %lld", (long long)v);
    if (g->row_off[0] != 0 || g->row_off[(size_t)V] != E) return fail(SMORE_E_INVALID, "row_off[0]/row_off[V] inconsistent with E");
    // degrees in the reference's accumulation order (src/proNet.cpp:423-446; pronet.go:198-212)
    g->n_local = V;
    g->out_deg.assign((size_t)V, 0.0);
    g->in_deg.assign((size_t)V, 0.0);
    std::vector<double>&out_deg = g->out_deg, &in_deg = g->in_deg;
    std::vector<double> dist((size_t)V);
    parallel_for(V, 1 << 14, [&](int64_t vb, int64_t ve, int) {  // per-vertex sums keep their in-order accumulation
        for (int64_t v = vb; v < ve; ++v)
            for (int64_t e = g->row_off[(size_t)v]; e < g->row_off[(size_t)v + 1]; ++e) out_deg[(size_t)v] += g->w[(size_t)e];
    });
    for (int64_t e = 0; e < E; ++e) in_deg[(size_t)g->col[(size_t)e]] += g->w[(size_t)e];  // order-dependent fp sum: serial
    if (g->sem == SMORE_SEM_CPP) {
        g->vertex_at = alias_method_cpp(out_deg.data(), V);  // vertex_method "out_degrees" (src/proNet.cpp:458-464)
        for (int64_t v = 0; v < V; ++v) {
            if (g->neg_method == SMORE_NEG_DEGREES) dist[(size_t)v] = in_deg[(size_t)v] + out_deg[(size_t)v];
            else if (g->neg_method == SMORE_NEG_IN_DEGREES) dist[(size_t)v] = in_deg[(size_t)v];
            else dist[(size_t)v] = in_deg[(size_t)v] == 0 ? 0 : 1;
        }
        g->negative_at = alias_method_cpp(dist.data(), V);
        g->ctx_at.prob.resize((size_t)E);
        g->ctx_at.alias.resize((size_t)E);
        // per-vertex sub-tables in CSR order (src/proNet.cpp:519-536): independent of each other -> one slice of vertices
        // per host thread, built in place (E pow() calls and E Vose steps: the bulk of the reference's start-up time)
        {
            const int nt = host_threads();
            const int64_t per = std::max<int64_t>(1, E / (int64_t)nt);  // slices balanced by CSR entries, not by vertices
            std::vector<int64_t> cut{0};
            for (int k = 1; k < nt; ++k) {
                const int64_t v = std::upper_bound(g->row_off.begin(), g->row_off.end(), per * k) - g->row_off.begin() - 1;
                cut.push_back(std::max(cut.back(), std::min(v, V)));
            }
            cut.push_back(V);
            parallel_for((int64_t)cut.size() - 1, 1, [&](int64_t sb, int64_t se, int) {
                AliasScratch sc;
                for (int64_t s = sb; s < se; ++s)
                    for (int64_t v = cut[(size_t)s]; v < cut[(size_t)s + 1]; ++v) {
                        const int64_t o = g->row_off[(size_t)v], b = g->row_off[(size_t)v + 1] - o;
                        if (b == 0) continue;
                        double* pr = g->ctx_at.prob.data() + o;
                        int64_t* al = g->ctx_at.alias.data() + o;
                        alias_method_cpp_into(g->w.data() + o, b, pr, al, sc);
                        for (int64_t i = 0; i < b; ++i)  // alias entries become vertex ids (src/proNet.cpp:530-534)
                            if (al[i] != -1) al[i] = (int64_t)g->col[(size_t)(o + al[i])];
                    }
            });
        }
    } else {
        g->vertex_at = alias_method_go(out_deg.data(), V, 1.0);  // pronet.go:224-230
        for (int64_t v = 0; v < V; ++v) dist[(size_t)v] = in_deg[(size_t)v] + out_deg[(size_t)v];
        g->negative_at = alias_method_go(dist.data(), V, 0.75);  // pronet.go:242-249
    }
    return upload_graph(g);
}


}  // namespace


namespace {
template <typename TH>
int rows_io_range(smore_model_t m, char* dptr, int64_t n, TH* host, bool to_device) {
    const int64_t cnt = n * (int64_t)m->dim;
    const bool same = (m->dtype == SMORE_F64) == (sizeof(TH) == 8);
    if (same) {
        if (to_device) CU(cudaMemcpy(dptr, host, (size_t)cnt * sizeof(TH), cudaMemcpyHostToDevice));
        else CU(cudaMemcpy((void*)host, dptr, (size_t)cnt * sizeof(TH), cudaMemcpyDeviceToHost));
        return SMORE_OK;
    }
    // dtype conversion on the device through a staging buffer
    TH* stage = nullptr;
    CU(cudaMalloc((void**)&stage, (size_t)cnt * sizeof(TH)));
    const int blocks = (int)std::min<int64_t>((cnt + 255) / 256, 148 * 16);
    int rc = SMORE_OK;
    if (to_device) {
        cudaError_t e = cudaMemcpy(stage, host, (size_t)cnt * sizeof(TH), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            if (m->dtype == SMORE_F64) k_convert<double, TH><<<blocks, 256>>>((double*)dptr, stage, cnt);
            else k_convert<float, TH><<<blocks, 256>>>((float*)dptr, stage, cnt);
            g_launches++;
            e = cudaDeviceSynchronize();
        }
        if (e != cudaSuccess) rc = fail(SMORE_E_CUDA, "set_rows: %s", cudaGetErrorString(e));
    } else {
        if (m->dtype == SMORE_F64) k_convert<TH, double><<<blocks, 256>>>(stage, (const double*)dptr, cnt);
        else k_convert<TH, float><<<blocks, 256>>>(stage, (const float*)dptr, cnt);
        g_launches++;
        cudaError_t e = cudaMemcpy((void*)host, stage, (size_t)cnt * sizeof(TH), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(SMORE_E_CUDA, "get_rows: %s", cudaGetErrorString(e));
    }
    cudaFree(stage);
    return rc;
}

template <typename TH>
int rows_io(smore_model_t m, int table, int64_t first, int64_t n, TH* host, bool to_device) {
    if (!m || table < 0 || table >= m->n_tables || !host) return fail(SMORE_E_INVALID, "bad model/table/buffer");
    if (first < 0 || n < 0 || first + n > m->rows) return fail(SMORE_E_INVALID, "row range out of bounds (the model holds %lld rows)", (long long)m->rows);
    if (n == 0) return SMORE_OK;
    if (int rc = ensure_device()) return rc;
    const int64_t first0 = first;
    return for_row_ranges(m, table, first, n, [&](char* dptr, int64_t f, int64_t k) {
        return rows_io_range<TH>(m, dptr, k, host + (size_t)(f - first0) * (size_t)m->dim, to_device);
    });
}
}  // namespace

// =================================================================================================================
extern "C" {

const char* smore_last_error(void) { return g_err.c_str(); }
const char* smore_version(void) { return "smore_b200 0.1 (sm_100a)"; }
uint64_t smore_kernel_launches(void) { return g_launches.load(); }

int smore_init(int device_id) {
    g_device = device_id;
    return ensure_device();
}

void smore_train_params_default(smore_train_params* p) {
    memset(p, 0, sizeof(*p));
    p->semantics = SMORE_SEM_CPP;
    p->mode = SMORE_MODE_HOGWILD;
    p->seed = 1;
    p->alpha = 0.025;
    p->total = 10ull * 1000000ull;
    p->negative_samples = 5;
    p->order = 2;
    p->lambda = 0.001;
    p->walk_times = 10;
    p->walk_steps = 40;
    p->window_min = 1;
    p->window_max = 5;
    p->max_walks = -1;
    p->n2v_p = 1.0;  // cmd/node2vec/main.go:21-22
    p->n2v_q = 1.0;
    p->item_reg = 0.01;    // cmd/cpr/main.go:22-24
    p->margin = 8.0;
    p->text_weight = 0.5;  // tpr.go:38
    p->xi = 10.0;  // cli/skewopt.cpp:53-54
    p->omega = 3.0;
    p->eta = 3;
}

int smore_graph_create(int64_t V, int64_t E, const int64_t* row_off, const int32_t* col, const double* weight,
                       int64_t n_lines, int semantics, int negative_method, smore_graph_t* out) {
    if (!row_off || (!col && E > 0) || (!weight && E > 0) || !out) return fail(SMORE_E_INVALID, "null argument");
    if (semantics != SMORE_SEM_CPP && semantics != SMORE_SEM_GO) return fail(SMORE_E_INVALID, "bad semantics");
    if (negative_method < 0 || negative_method > 2) return fail(SMORE_E_INVALID, "bad negative_method");
    // validate the shape BEFORE touching the caller's arrays: ids are int32 and index draws are 32-bit on the device
    if (V <= 0 || V >= (1ll << 31)) return fail(SMORE_E_INVALID, "V=%lld out of range (0, 2^31)", (long long)V);
    if (E < 0 || E >= (1ll << 32)) return fail(SMORE_E_INVALID, "E=%lld out of range [0, 2^32)", (long long)E);
    if (row_off[0] != 0 || row_off[V] != E) return fail(SMORE_E_INVALID, "row_off[0]/row_off[V] inconsistent with E");
    if (int rc = ensure_device()) return rc;
    smore_graph_s* g = new smore_graph_s();
    g->sem = semantics;
    g->neg_method = semantics == SMORE_SEM_GO ? SMORE_NEG_DEGREES : negative_method;
    g->V = V;
    g->E = E;
    g->n_lines = n_lines > 0 ? n_lines : E;
    g->row_off.assign(row_off, row_off + V + 1);
    g->col.assign(col, col + E);
    g->w.assign(weight, weight + E);
    int rc = build_graph(g);
    if (rc) { delete g; return rc; }
    *out = g;
    return SMORE_OK;
}

int smore_graph_load_edge_list(const char* path, int undirected, int semantics, int negative_method, smore_graph_t* out) {
    if (!path || !out) return fail(SMORE_E_INVALID, "null argument");
    if (semantics != SMORE_SEM_CPP && semantics != SMORE_SEM_GO) return fail(SMORE_E_INVALID, "bad semantics");
    if (int rc = ensure_device()) return rc;
    EdgeList el;
    std::string err;
    if (!load_edge_list(path, undirected != 0, el, err)) return fail(SMORE_E_IO, "%s", err.c_str());
    if (el.names.empty()) return fail(SMORE_E_IO, "no edges in %s", path);
    smore_graph_s* g = new smore_graph_s();
    g->sem = semantics;
    g->neg_method = semantics == SMORE_SEM_GO ? SMORE_NEG_DEGREES : negative_method;
    g->V = (int64_t)el.names.size();
    g->E = (int64_t)el.col.size();
    // C++: MAX_line counts CSR entries (doubled when undirected, src/proNet.cpp:230-231); Go: edge lines (pronet.go:164)
    g->n_lines = semantics == SMORE_SEM_CPP ? g->E : el.n_lines;
    g->row_off = std::move(el.row_off);
    g->col = std::move(el.col);
    g->w = std::move(el.w);
    g->names = std::move(el.names);
    int rc = build_graph(g);
    if (rc) { delete g; return rc; }
    *out = g;
    return SMORE_OK;
}

int smore_edge_list_to_csr(const char* path, int undirected, int64_t* V, int64_t* E, int64_t* n_lines, int64_t* row_off,
                           int32_t* col, double* weight, char* names, int64_t names_cap, int64_t* names_bytes) {
    if (!path) return fail(SMORE_E_INVALID, "null path");
    EdgeList el;
    std::string err;
    if (!load_edge_list(path, undirected != 0, el, err)) return fail(SMORE_E_IO, "%s", err.c_str());
    const int64_t nv = (int64_t)el.names.size(), ne = (int64_t)el.col.size();
    if (V) *V = nv;
    if (E) *E = ne;
    if (n_lines) *n_lines = el.n_lines;
    if (row_off) memcpy(row_off, el.row_off.data(), (size_t)(nv + 1) * sizeof(int64_t));
    if (col) memcpy(col, el.col.data(), (size_t)ne * sizeof(int32_t));
    if (weight) memcpy(weight, el.w.data(), (size_t)ne * sizeof(double));
    int64_t nb = 0;
    for (auto& s : el.names) nb += (int64_t)s.size() + 1;
    if (names_bytes) *names_bytes = nb;
    if (names) {
        if (names_cap < nb) return fail(SMORE_E_INVALID, "names buffer too small (%lld < %lld)", (long long)names_cap, (long long)nb);
        char* p = names;
        for (auto& s : el.names) {
            memcpy(p, s.data(), s.size());
            p += s.size();
            *p++ = '\n';
        }
    }
    return SMORE_OK;
}

int smore_graph_set_field(smore_graph_t g, const int32_t* field) {
    if (!g || !field) return fail(SMORE_E_INVALID, "null argument");
    if (g->synthetic) return fail(SMORE_E_UNSUPPORTED, "this graph was generated on the device (no host CSR / reference tables): only the rotating-shard trainer works on it");
    g->field.assign(field, field + g->V);
    g->has_field = true;
    CU(cudaMemcpy(g->d_field, g->field.data(), (size_t)g->V * sizeof(int32_t), cudaMemcpyHostToDevice));
    return SMORE_OK;
}

int smore_graph_load_field(smore_graph_t g, const char* path) {
    if (!g || !path) return fail(SMORE_E_INVALID, "null argument");
    if (g->names.empty()) return fail(SMORE_E_INVALID, "graph has no vertex names (created from CSR); use smore_graph_set_field");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(SMORE_E_IO, "cannot open field file: %s", path);
    std::map<std::string, int32_t> name_to_id;
    for (size_t i = 0; i < g->names.size(); ++i) name_to_id.emplace(g->names[i], (int32_t)i);
    std::map<std::string, int32_t> meta_idx;  // field ids in first-appearance order (src/proNet.cpp:368-372)
    std::vector<int32_t> field((size_t)g->V, 0);
    char v[1024], meta[1024];
    while (fscanf(f, "%1023s %1023s", v, meta) == 2) {
        auto it = meta_idx.find(meta);
        if (it == meta_idx.end()) it = meta_idx.emplace(meta, (int32_t)meta_idx.size()).first;
        auto vit = name_to_id.find(v);
        if (vit != name_to_id.end()) field[(size_t)vit->second] = it->second;
    }
    fclose(f);
    return smore_graph_set_field(g, field.data());
}

int smore_graph_info(smore_graph_t g, int64_t* V, int64_t* E, int64_t* n_lines) {
    if (!g) return fail(SMORE_E_INVALID, "null graph");
    if (V) *V = g->V;
    if (E) *E = g->E;
    if (n_lines) *n_lines = g->n_lines;
    return SMORE_OK;
}

int smore_graph_get_csr(smore_graph_t g, int64_t* row_off, int32_t* col, double* weight) {
    if (!g) return fail(SMORE_E_INVALID, "null graph");
    if (g->synthetic) return fail(SMORE_E_UNSUPPORTED, "this graph was generated on the device (no host CSR / reference tables): only the rotating-shard trainer works on it");
    if (row_off) memcpy(row_off, g->row_off.data(), g->row_off.size() * sizeof(int64_t));
    if (col) memcpy(col, g->col.data(), g->col.size() * sizeof(int32_t));
    if (weight) memcpy(weight, g->w.data(), g->w.size() * sizeof(double));
    return SMORE_OK;
}

const char* smore_graph_vertex_name(smore_graph_t g, int64_t vid) {
    if (!g || vid < 0 || vid >= (int64_t)g->names.size()) return nullptr;
    return g->names[(size_t)vid].c_str();
}

int smore_graph_get_alias(smore_graph_t g, int which, double* prob, int64_t* alias) {
    if (!g) return fail(SMORE_E_INVALID, "null graph");
    if (g->synthetic) return fail(SMORE_E_UNSUPPORTED, "this graph was generated on the device (no host CSR / reference tables): only the rotating-shard trainer works on it");
    const AliasHost* t = which == SMORE_AT_VERTEX ? &g->vertex_at : which == SMORE_AT_NEGATIVE ? &g->negative_at : which == SMORE_AT_CONTEXT ? &g->ctx_at : nullptr;
    if (!t) return fail(SMORE_E_INVALID, "bad table selector");
    if (which == SMORE_AT_CONTEXT && g->sem != SMORE_SEM_CPP) return fail(SMORE_E_UNSUPPORTED, "context alias table exists only under C++ semantics");
    if (prob) memcpy(prob, t->prob.data(), t->prob.size() * sizeof(double));
    if (alias) memcpy(alias, t->alias.data(), t->alias.size() * sizeof(int64_t));
    return SMORE_OK;
}

int smore_graph_get_field(smore_graph_t g, int32_t* field) {
    if (!g || !field) return fail(SMORE_E_INVALID, "null argument");
    if (g->synthetic) return fail(SMORE_E_UNSUPPORTED, "this graph was generated on the device (no host CSR / reference tables): only the rotating-shard trainer works on it");
    memcpy(field, g->field.data(), g->field.size() * sizeof(int32_t));
    return SMORE_OK;
}

void smore_graph_destroy(smore_graph_t g) { delete g; }

int smore_sample_debug(smore_graph_t g, int which, uint64_t seed, uint64_t stream, int64_t n, const int64_t* arg,
                       int64_t* out, uint64_t* words_used) {
    if (!g || !out || n < 0) return fail(SMORE_E_INVALID, "bad argument");
    if (g->synthetic) return fail(SMORE_E_UNSUPPORTED, "this graph was generated on the device (no host CSR / reference tables): only the rotating-shard trainer works on it");
    if (which < 0 || which > 3) return fail(SMORE_E_INVALID, "bad sampler selector");
    if (which == SMORE_SAMPLE_TARGET && !arg) return fail(SMORE_E_INVALID, "TARGET needs source vertices");
    if (int rc = ensure_device()) return rc;
    const int64_t n_out = which == 3 ? 2 * n : n;
    int64_t *d_arg = nullptr, *d_out = nullptr;
    uint64_t* d_used = nullptr;
    if (arg) {
        for (int64_t i = 0; i < n; ++i)
            if (arg[i] < 0 || arg[i] >= g->V) return fail(SMORE_E_INVALID, "arg[%lld] out of range", (long long)i);
        CU(cudaMalloc((void**)&d_arg, (size_t)std::max<int64_t>(n, 1) * 8));
        CU(cudaMemcpy(d_arg, arg, (size_t)n * 8, cudaMemcpyHostToDevice));
    }
    CU(cudaMalloc((void**)&d_out, (size_t)std::max<int64_t>(n_out, 1) * 8));
    CU(cudaMalloc((void**)&d_used, 8));
    k_sample_debug<<<1, 32>>>(g->view(), which, seed, stream, n, d_arg, d_out, d_used);
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, d_out, (size_t)n_out * 8, cudaMemcpyDeviceToHost));
    if (words_used) CU(cudaMemcpy(words_used, d_used, 8, cudaMemcpyDeviceToHost));
    cudaFree(d_arg); cudaFree(d_out); cudaFree(d_used);
    return SMORE_OK;
}

int smore_walk_debug(smore_graph_t g, uint64_t seed, uint64_t stream, int64_t start, int steps, int mode, int w0, int w1,
                     int64_t* walk, int64_t* walk_len, int64_t* pair_v, int64_t* pair_c, int64_t cap, int64_t* n_pairs) {
    if (!g || !walk || !walk_len || !n_pairs) return fail(SMORE_E_INVALID, "null argument");
    if (g->synthetic) return fail(SMORE_E_UNSUPPORTED, "this graph was generated on the device (no host CSR / reference tables): only the rotating-shard trainer works on it");
    if (start < 0 || start >= g->V) return fail(SMORE_E_INVALID, "start out of range");
    if (steps < 0 || steps + 1 > kMaxWalkLen) return fail(SMORE_E_UNSUPPORTED, "walk_steps must be < %d", kMaxWalkLen);
    if (mode == 0 && (w0 < 1 || w0 > 255)) return fail(SMORE_E_INVALID, "window must be in [1,255]");
    if (mode == 1 && g->sem != SMORE_SEM_CPP) return fail(SMORE_E_UNSUPPORTED, "ScaleSkipGrams exists only under C++ semantics");
    if (int rc = ensure_device()) return rc;
    int64_t *d_walk, *d_len, *d_pv, *d_pc, *d_np;
    const int64_t c = std::max<int64_t>(cap, 1);
    CU(cudaMalloc((void**)&d_walk, (size_t)(steps + 1) * 8));
    CU(cudaMalloc((void**)&d_len, 8));
    CU(cudaMalloc((void**)&d_np, 8));
    CU(cudaMalloc((void**)&d_pv, (size_t)c * 8));
    CU(cudaMalloc((void**)&d_pc, (size_t)c * 8));
    k_walk_debug<<<1, 32>>>(g->view(), seed, stream, start, steps, mode, w0, w1, d_walk, d_len, d_pv, d_pc, cap, d_np);
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpy(walk_len, d_len, 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(n_pairs, d_np, 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(walk, d_walk, (size_t)(*walk_len) * 8, cudaMemcpyDeviceToHost));
    const int64_t ncopy = std::min<int64_t>(*n_pairs, cap);
    if (pair_v && ncopy > 0) CU(cudaMemcpy(pair_v, d_pv, (size_t)ncopy * 8, cudaMemcpyDeviceToHost));
    if (pair_c && ncopy > 0) CU(cudaMemcpy(pair_c, d_pc, (size_t)ncopy * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_walk); cudaFree(d_len); cudaFree(d_np); cudaFree(d_pv); cudaFree(d_pc);
    return SMORE_OK;
}

// ---- embedding store ----------------------------------------------------------------------------------------------
int smore_model_create(smore_graph_t g, int dim, int n_tables, int dtype, smore_model_t* out) {
    if (!g || !out) return fail(SMORE_E_INVALID, "null argument");
    if (n_tables < 1 || n_tables > 2) return fail(SMORE_E_INVALID, "n_tables must be 1 or 2");
    if (dtype != SMORE_F32 && dtype != SMORE_F64) return fail(SMORE_E_INVALID, "bad dtype");
    if (dim <= 0) return fail(SMORE_E_INVALID, "dim must be positive");
    if (int rc = ensure_device()) return rc;
    smore_model_s* m = new smore_model_s();
    m->g = g;
    m->dim = dim;
    m->n_tables = n_tables;
    m->dtype = dtype;
    m->rows = g->n_local;  // == V unless the graph is row-sharded
    for (int t = 0; t < n_tables; ++t) {
        size_t bytes = (size_t)std::max<int64_t>(m->rows, 1) * (size_t)dim * m->elem();
        cudaError_t e = cudaMalloc(&m->tab[t], bytes);
        if (e != cudaSuccess) {
            delete m;
            return fail(SMORE_E_NOMEM, "cudaMalloc of %zu bytes for table %d failed: %s", bytes, t, cudaGetErrorString(e));
        }
        cudaMemset(m->tab[t], 0, bytes);
        m->peer[t][g->rank] = m->tab[t];
    }
    g->n_models++;
    *out = m;
    return SMORE_OK;
}

// ---- row sharding across the GPUs of one box (one process per GPU) ---------------------------------------------------
int smore_graph_set_shard(smore_graph_t g, int rank, int world) {
    if (!g) return fail(SMORE_E_INVALID, "null graph");
    if (g->synthetic) return fail(SMORE_E_UNSUPPORTED, "this graph was generated on the device (no host CSR / reference tables): only the rotating-shard trainer works on it");
    if (world < 1 || world > kMaxWorld || (world & (world - 1))) return fail(SMORE_E_INVALID, "world must be 1, 2, 4 or 8");
    if (rank < 0 || rank >= world) return fail(SMORE_E_INVALID, "rank out of range");
    if (g->n_models > 0) return fail(SMORE_E_INVALID, "%d model(s) were already created on this graph: their row counts are fixed at creation, shard the graph first", g->n_models);
    if (int rc = ensure_device()) return rc;
    int shift = 0;
    while ((1 << shift) < world) ++shift;
    const int64_t V = g->V;
    const int64_t nl = (V - rank + world - 1) / world;
    if (nl <= 0) return fail(SMORE_E_INVALID, "rank %d owns no vertex", rank);
    if (world == 1) return SMORE_OK;  // nothing to shard: the reference's own samplers stay in place
    // The rank that owns the positive context computes the sample. Edge table: every CSR entry (v1 -> v2) with v2 owned
    // here, weighted by the probability the unsharded samplers give it: P(source = v1) * P(target = v2 | v1)
    //   C++: out_deg(v1)^0.75 / sum  *  w_e^0.75 / sum_{e' of v1} w_e'^0.75   (AliasMethod applies 0.75 to every table)
    //   Go : out_deg(v1) / sum       *  w_e / out_deg(v1)
    // so the union of the ranks' samples, rank r running total * mass_r of them, is the global edge distribution.
    // Negative table: the owned vertices only (local index l <-> vertex l*world + rank), same construction as the global
    // one: negatives are stratified by the shard of the positive.
    const double pw = g->sem == SMORE_SEM_CPP ? 0.75 : 1.0;
    std::vector<double> psrc((size_t)V), nrm((size_t)V, 0.0);
    double src_sum = 0;
    auto wpow = [pw](double x) { return pw == 0.75 ? pow075(x) : x; };  // pw is 0.75 (C++) or 1 (Go)
    parallel_for(V, 1 << 14, [&](int64_t vb, int64_t ve, int) {
        for (int64_t v = vb; v < ve; ++v) {
            psrc[(size_t)v] = g->out_deg[(size_t)v] > 0 ? std::pow(g->out_deg[(size_t)v], pw) : 0.0;
            for (int64_t e = g->row_off[(size_t)v]; e < g->row_off[(size_t)v + 1]; ++e) nrm[(size_t)v] += wpow(g->w[(size_t)e]);
        }
    });
    for (int64_t v = 0; v < V; ++v) src_sum += psrc[(size_t)v];
    std::vector<double> pe;
    std::vector<int32_t> esrc, edst;
    double mass_all = 0, mass_own = 0;
    for (int64_t v = 0; v < V; ++v) {
        if (nrm[(size_t)v] <= 0 || src_sum <= 0) continue;
        const double pv = psrc[(size_t)v] / src_sum / nrm[(size_t)v];
        for (int64_t e = g->row_off[(size_t)v]; e < g->row_off[(size_t)v + 1]; ++e) {
            const double pr = pv * wpow(g->w[(size_t)e]);
            mass_all += pr;
            if ((g->col[(size_t)e] & (world - 1)) == rank) {
                mass_own += pr;
                pe.push_back(pr);
                esrc.push_back((int32_t)v);
                edst.push_back(g->col[(size_t)e]);
            }
        }
    }
    const int64_t ne = (int64_t)pe.size();
    if (ne == 0) return fail(SMORE_E_INVALID, "rank %d owns no edge target", rank);
    if (ne >= (1ll << 32)) return fail(SMORE_E_UNSUPPORTED, "rank %d owns %lld CSR entries; the edge table is indexed with 32-bit draws", rank, (long long)ne);
    std::vector<double> neg((size_t)nl);
    for (int64_t l = 0; l < nl; ++l) {
        const int64_t v = l * world + rank;
        const double in = g->in_deg[(size_t)v], out = g->out_deg[(size_t)v];
        if (g->sem == SMORE_SEM_GO || g->neg_method == SMORE_NEG_DEGREES) neg[(size_t)l] = in + out;
        else if (g->neg_method == SMORE_NEG_IN_DEGREES) neg[(size_t)l] = in;
        else neg[(size_t)l] = in == 0 ? 0 : 1;
    }
    AliasHost edge_at = alias_method_go(pe.data(), ne, 1.0);  // plain Vose on the edge probabilities
    // A/B switch (tools/ab_sharded_quality.py): keep the table over ALL vertices -- negative context rows then live on any
    // rank and cross NVLink like the vertex rows do ("global negatives with P2P remote rows")
    const char* gn = getenv("SMORE_SHARD_GLOBAL_NEG");
    g->neg_global = gn && atoi(gn) != 0;
    g->n_neg = g->neg_global ? V : nl;
    if (!g->neg_global)
        g->negative_at = g->sem == SMORE_SEM_CPP ? alias_method_cpp(neg.data(), nl) : alias_method_go(neg.data(), nl, 0.75);
    std::vector<uint2> packed;
    auto upload = [&](const AliasHost& t, uint2** d) -> int {
        packed.resize(t.prob.size());
        for (size_t i = 0; i < t.prob.size(); ++i) {
            PackedAlias pa = pack_alias(t.prob[i], t.alias[i], (uint32_t)i);
            packed[i] = make_uint2(pa.thr, pa.alias);
        }
        cudaFree(*d);
        return dev_alloc_copy(d, packed.data(), packed.size());
    };
    if (int rc = upload(edge_at, &g->d_eat)) return rc;
    if (!g->neg_global)
        if (int rc = upload(g->negative_at, &g->d_nat)) return rc;
    cudaFree(g->d_esrc);
    cudaFree(g->d_edst);
    if (int rc = dev_alloc_copy(&g->d_esrc, esrc.data(), esrc.size())) return rc;
    if (int rc = dev_alloc_copy(&g->d_edst, edst.data(), edst.size())) return rc;
    g->n_edge_local = ne;
    g->rank = rank;
    g->world = world;
    g->shift = shift;
    g->n_local = nl;
    g->src_mass_frac = mass_all > 0 ? mass_own / mass_all : 1.0 / world;
    return SMORE_OK;
}

int smore_graph_shard_info(smore_graph_t g, int* rank, int* world, int64_t* n_local, double* source_mass_fraction) {
    if (!g) return fail(SMORE_E_INVALID, "null graph");
    if (rank) *rank = g->rank;
    if (world) *world = g->world;
    if (n_local) *n_local = g->n_local;
    if (source_mass_fraction) *source_mass_fraction = g->src_mass_frac;
    return SMORE_OK;
}

int smore_model_ipc_handle(smore_model_t m, int table, void* handle64) {
    if (!m || table < 0 || table >= m->n_tables || !handle64) return fail(SMORE_E_INVALID, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!m->tab[table]) return fail(SMORE_E_INVALID, "table %d rotates between ranks (smore_model_rot_ipc_handles)", table);
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, m->tab[table]));
    memcpy(handle64, &h, 64);
    return SMORE_OK;
}

int smore_model_open_peers(smore_model_t m, int table, const void* handles) {
    if (!m || table < 0 || table >= m->n_tables || !handles) return fail(SMORE_E_INVALID, "bad argument");
    if (int rc = ensure_device()) return rc;
    for (int r = 0; r < m->g->world; ++r) {
        if (r == m->g->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * r, 64);
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        m->peer[table][r] = p;
        m->peer_opened[table][r] = true;
    }
    return SMORE_OK;
}

int smore_model_enable_replica(smore_model_t m, int table) {
    if (!m || table < 0 || table >= m->n_tables) return fail(SMORE_E_INVALID, "bad model/table");
    if (m->g->world == 1) return fail(SMORE_E_INVALID, "replicas only make sense on a row-sharded graph");
    if (m->g->rotating) return fail(SMORE_E_INVALID, "the graph is set up for rotating shards");
    if (((size_t)m->dim * m->elem()) % 16) return fail(SMORE_E_UNSUPPORTED, "replica rows must be a multiple of 16 bytes");
    if (int rc = ensure_device()) return rc;
    if (!m->replica[table]) {
        size_t bytes = (size_t)m->g->V * (size_t)m->dim * m->elem();
        cudaError_t e = cudaMalloc(&m->replica[table], bytes);
        if (e != cudaSuccess) return fail(SMORE_E_NOMEM, "cudaMalloc of %zu bytes for the replica failed: %s", bytes, cudaGetErrorString(e));
    }
    return SMORE_OK;
}

int smore_model_refresh_replica(smore_model_t m, int table) {
    if (!m || table < 0 || table >= m->n_tables || !m->replica[table]) return fail(SMORE_E_INVALID, "no replica enabled for this table");
    for (int r = 0; r < m->g->world; ++r)
        if (!m->peer[table][r]) return fail(SMORE_E_INVALID, "shard of rank %d not connected", r);
    if (int rc = ensure_device()) return rc;
    const int64_t n_vec = m->g->V * (((int64_t)m->dim * (int64_t)m->elem()) / 16);
    const int blocks = (int)std::min<int64_t>((n_vec + 255) / 256, 148 * 16);
    if (m->dtype == SMORE_F64) {
        PeerBases<double> pb;
        for (int r = 0; r < kMaxWorld; ++r) pb.b[r] = (double*)m->peer[table][r];
        k_refresh_replica<double><<<blocks, 256>>>((double*)m->replica[table], pb, m->g->shift, m->g->world - 1, m->dim, m->g->V);
    } else {
        PeerBases<float> pb;
        for (int r = 0; r < kMaxWorld; ++r) pb.b[r] = (float*)m->peer[table][r];
        k_refresh_replica<float><<<blocks, 256>>>((float*)m->replica[table], pb, m->g->shift, m->g->world - 1, m->dim, m->g->V);
    }
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    return SMORE_OK;
}

int smore_model_set_peer_ptrs(smore_model_t m, int table, void* const* ptrs) {
    if (!m || table < 0 || table >= m->n_tables || !ptrs) return fail(SMORE_E_INVALID, "bad argument");
    for (int r = 0; r < m->g->world; ++r)
        if (r != m->g->rank) m->peer[table][r] = ptrs[r];
    return SMORE_OK;
}

void smore_model_destroy(smore_model_t m) { delete m; }

int smore_model_init(smore_model_t m, int table, int random, uint64_t seed) {
    if (!m || table < 0 || table >= m->n_tables) return fail(SMORE_E_INVALID, "bad model/table");
    if (int rc = ensure_device()) return rc;
    const int sh = m->g->shift, rk = m->g->rank;  // a shard initialises its rows with the words of their GLOBAL position
    if (int rc = for_row_ranges(m, table, 0, m->rows, [&](char* dptr, int64_t first, int64_t k) -> int {
            const int64_t n = k * (int64_t)m->dim;
            const int blocks = (int)std::min<int64_t>((n + 255) / 256 + 1, 148 * 16);
            if (m->dtype == SMORE_F64) k_init_table<double><<<blocks, 256>>>((double*)dptr, n, m->dim, seed, kInitStreamBase + (uint64_t)table, random, sh, rk, first);
            else k_init_table<float><<<blocks, 256>>>((float*)dptr, n, m->dim, seed, kInitStreamBase + (uint64_t)table, random, sh, rk, first);
            g_launches++;
            CU(cudaGetLastError());
            return SMORE_OK;
        }))
        return rc;
    CU(cudaDeviceSynchronize());
    return SMORE_OK;
}


int smore_model_set_rows(smore_model_t m, int table, int64_t first, int64_t n, const double* host) {
    return rows_io<double>(m, table, first, n, const_cast<double*>(host), true);
}
int smore_model_get_rows(smore_model_t m, int table, int64_t first, int64_t n, double* host) {
    return rows_io<double>(m, table, first, n, host, false);
}
int smore_model_set_rows_f32(smore_model_t m, int table, int64_t first, int64_t n, const float* host) {
    return rows_io<float>(m, table, first, n, const_cast<float*>(host), true);
}
int smore_model_get_rows_f32(smore_model_t m, int table, int64_t first, int64_t n, float* host) {
    return rows_io<float>(m, table, first, n, host, false);
}

// Asynchronous row transfers for hosts that pipeline their own staging (bench.py's end-to-end leg: the read-back of step i
// overlaps the upload of step i+1 -- PCIe is full duplex). fp32 models only, no dtype conversion; `host` should be pinned
// (page-locked) memory, otherwise the copies degrade to synchronous ones. Uploads and read-backs run on two internal
// streams that do not synchronise with the trainers: call smore_model_wait_copies before training on rows that are being
// uploaded, and before touching host memory that is being filled.
static int async_rows(smore_model_t m, int table, int64_t first, int64_t n, float* host, bool to_device) {
    if (!m || table < 0 || table >= m->n_tables || !host) return fail(SMORE_E_INVALID, "bad model/table/buffer");
    if (m->dtype != SMORE_F32) return fail(SMORE_E_UNSUPPORTED, "asynchronous row transfers exist for fp32 models only");
    if (first < 0 || n < 0 || first + n > m->rows) return fail(SMORE_E_INVALID, "row range out of bounds (the model holds %lld rows)", (long long)m->rows);
    if (n == 0) return SMORE_OK;
    if (int rc = ensure_device()) return rc;
    cudaStream_t& st = to_device ? m->h2d_stream : m->d2h_stream;
    if (!st) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    // Device-side ordering between the two copy streams of ONE model, per TABLE: an upload is ordered after the read-backs of
    // the same table issued before it (it may overwrite the rows they are still reading) and a read-back after the uploads of
    // that table issued before it -- so a host that double-buffers two models never has to block on a copy before it can
    // launch the next train call, and the upload of one table runs next to the read-back of the other (PCIe is full duplex).
    cudaEvent_t& mine = to_device ? m->h2d_event[table] : m->d2h_event[table];
    cudaEvent_t& other = to_device ? m->d2h_event[table] : m->h2d_event[table];
    if (!mine) CU(cudaEventCreateWithFlags(&mine, cudaEventDisableTiming));
    if (other) CU(cudaStreamWaitEvent(st, other, 0));
    const size_t row_bytes = (size_t)m->dim * sizeof(float);
    const int64_t first0 = first;
    return for_row_ranges(m, table, first, n, [&](char* dptr, int64_t f, int64_t k) -> int {
        char* h = (char*)host + (size_t)(f - first0) * row_bytes;
        if (to_device) CU(cudaMemcpyAsync(dptr, h, (size_t)k * row_bytes, cudaMemcpyHostToDevice, st));
        else CU(cudaMemcpyAsync(h, dptr, (size_t)k * row_bytes, cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(mine, st));
        return SMORE_OK;
    });
}
int smore_model_set_rows_f32_async(smore_model_t m, int table, int64_t first, int64_t n, const float* host) {
    return async_rows(m, table, first, n, const_cast<float*>(host), true);
}
int smore_model_get_rows_f32_async(smore_model_t m, int table, int64_t first, int64_t n, float* host) {
    return async_rows(m, table, first, n, host, false);
}
int smore_model_wait_copies(smore_model_t m, int uploads, int readbacks) {
    if (!m) return fail(SMORE_E_INVALID, "null model");
    if (uploads && m->h2d_stream) CU(cudaStreamSynchronize(m->h2d_stream));
    if (readbacks && m->d2h_stream) CU(cudaStreamSynchronize(m->d2h_stream));
    return SMORE_OK;
}

int smore_model_device_ptr(smore_model_t m, int table, void** ptr) {
    if (!m || table < 0 || table >= m->n_tables || !ptr) return fail(SMORE_E_INVALID, "bad model/table");
    if (!m->tab[table]) return fail(SMORE_E_INVALID, "table %d rotates between ranks: it has no fixed device address (smore_model_rot_slot_ptrs)", table);
    *ptr = m->tab[table];
    return SMORE_OK;
}

int smore_model_save_weights(smore_model_t m, int table, const char* path, int format) {
    if (!m || !path || table < 0 || table >= m->n_tables) return fail(SMORE_E_INVALID, "bad argument");
    if (m->g->world > 1) return fail(SMORE_E_UNSUPPORTED, "save_weights on a row-sharded model: gather the shards with get_rows");
    const int64_t V = m->g->V;
    const int dim = m->dim;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(SMORE_E_IO, "cannot create %s", path);
    fprintf(f, "%lld %d\n", (long long)V, dim);
    // 64 MB of rows at a time: device -> host, then every host thread formats its slice of the chunk into its own buffer
    // (std::to_chars: the reference's number formats without the printf machinery) and the buffers are written in order
    const int64_t chunk = std::max<int64_t>(1, (64ll << 20) / ((int64_t)dim * 8));
    std::vector<double> rows((size_t)chunk * (size_t)dim);
    std::vector<std::string> parts((size_t)host_threads());
    const std::string* names = m->g->names.empty() ? nullptr : m->g->names.data();
    for (int64_t first = 0; first < V; first += chunk) {
        const int64_t n = std::min(chunk, V - first);
        int rc = smore_model_get_rows(m, table, first, n, rows.data());
        if (rc) { fclose(f); return rc; }
        for (auto& s : parts) s.clear();
        parallel_for(n, 256, [&](int64_t b, int64_t e, int t) {
            parts[(size_t)t].reserve((size_t)(e - b) * (size_t)dim * 10);
            format_rows(rows.data() + (size_t)b * (size_t)dim, e - b, dim, names, first + b, format, parts[(size_t)t]);
        });
        for (auto& s : parts)
            if (!s.empty() && fwrite(s.data(), 1, s.size(), f) != s.size()) { fclose(f); return fail(SMORE_E_IO, "short write to %s", path); }
    }
    fclose(f);
    return SMORE_OK;
}

// ---- warm start and checkpoints --------------------------------------------------------------------------------------
// proNet::LoadPreTrain (src/proNet.cpp:238-286; cli/deepwalk.cpp:61-62 -load_v / -load_c): "<n> <dim>" then `name v0 v1 ...`
// per line; rows whose name the graph does not know are skipped, a dimension mismatch skips the whole file. (The reference
// reads lines with fgets(…, 1000) and silently truncates longer ones, i.e. dim > ~90; that defect is not reproduced.)
int smore_model_load_pretrain(smore_model_t m, int table, const char* path, int64_t* rows_loaded) {
    if (!m || !path || table < 0 || table >= m->n_tables) return fail(SMORE_E_INVALID, "bad argument");
    if (rows_loaded) *rows_loaded = 0;
    const smore_graph_s* g = m->g;
    if (g->names.empty()) return fail(SMORE_E_INVALID, "the graph has no vertex names (created from a CSR): nothing to match the file against");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(SMORE_E_IO, "cannot open %s", path);
    std::string buf;
    {
        fseek(f, 0, SEEK_END);
        const long size = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf.resize((size_t)std::max(0l, size));
        if (size > 0 && fread(&buf[0], 1, (size_t)size, f) != (size_t)size) {
            fclose(f);
            return fail(SMORE_E_IO, "short read: %s", path);
        }
        fclose(f);
    }
    const char* p = buf.data();
    const char* end = p + buf.size();
    auto next_line = [&](const char*& b, const char*& e) -> bool {
        if (p >= end) return false;
        b = p;
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        e = nl ? nl : end;
        p = nl ? nl + 1 : end;
        return true;
    };
    auto next_tok = [](const char*& q, const char* e, const char*& tb, const char*& te) -> bool {
        while (q < e && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
        if (q >= e) return false;
        tb = q;
        while (q < e && !(*q == ' ' || *q == '\t' || *q == '\r')) ++q;
        te = q;
        return true;
    };
    const char *lb, *le, *tb, *te;
    if (!next_line(lb, le)) return SMORE_OK;  // empty file: nothing loaded (the reference returns silently too)
    long long n_hdr = 0, dim_hdr = 0;
    {
        const char* q = lb;
        if (next_tok(q, le, tb, te)) n_hdr = atoll(std::string(tb, te).c_str());
        if (next_tok(q, le, tb, te)) dim_hdr = atoll(std::string(tb, te).c_str());
    }
    (void)n_hdr;
    if (dim_hdr != m->dim) return SMORE_OK;  // "Dimension not matched, Skip Loading Pre-train model." (proNet.cpp:262-264)
    std::unordered_map<std::string_view, int64_t> ids;
    ids.reserve(g->names.size() * 2);
    for (size_t v = 0; v < g->names.size(); ++v) ids.emplace(std::string_view(g->names[v]), (int64_t)v);
    // consecutive local rows are flushed as one range (a file written by save_weights is in id order: one copy per 16 MB)
    const int64_t cap_rows = std::max<int64_t>(1, (16ll << 20) / ((int64_t)m->dim * 8));
    std::vector<double> run;
    run.reserve((size_t)cap_rows * (size_t)m->dim);
    int64_t run_first = -1, run_n = 0, loaded = 0;
    auto flush = [&]() -> int {
        int rc = SMORE_OK;
        if (run_n > 0) rc = smore_model_set_rows(m, table, run_first, run_n, run.data());
        run.clear();
        run_first = -1;
        run_n = 0;
        return rc;
    };
    std::vector<double> row((size_t)m->dim);
    while (next_line(lb, le)) {
        const char* q = lb;
        if (!next_tok(q, le, tb, te)) continue;
        auto it = ids.find(std::string_view(tb, (size_t)(te - tb)));
        if (it == ids.end()) continue;
        int d = 0;
        while (d < m->dim && next_tok(q, le, tb, te)) {
            double x = 0;
            auto r = std::from_chars(tb, te, x);
            if (r.ec != std::errc()) x = atof(std::string(tb, te).c_str());
            row[(size_t)d++] = x;
        }
        if (d != m->dim) continue;  // short line: skipped
        const int64_t v = it->second;
        if ((v & (g->world - 1)) != g->rank) continue;  // row-sharded: another rank's row
        const int64_t local = v >> g->shift;
        if (run_n > 0 && (local != run_first + run_n || run_n >= cap_rows))
            if (int rc = flush()) return rc;
        if (run_n == 0) run_first = local;
        run.insert(run.end(), row.begin(), row.end());
        ++run_n;
        ++loaded;
    }
    if (int rc = flush()) return rc;
    if (rows_loaded) *rows_loaded = loaded;
    return SMORE_OK;
}

namespace {
struct CkptHeader {
    char magic[8];  // "SMOREB2\0"
    uint32_t version, dtype, n_tables, dim;
    int64_t V, rows;
    int32_t rank, world;
};
// version 2 appends the training position, so that a checkpoint is a RESUME point and not only a warm start: the Philox
// key, the first sampler stream no call has used yet, and where the LR schedule stands (units of `total`)
struct CkptProgress {
    uint64_t seed, next_stream, sched_total, sched_done;
};
constexpr char kCkptMagic[8] = {'S', 'M', 'O', 'R', 'E', 'B', '2', '\0'};
}  // namespace

// Binary snapshot of this rank's rows of every table (raw, in the stored element type): the reference can import text
// embeddings but cannot resume; a sharded run writes one file per rank.
int smore_model_save_checkpoint(smore_model_t m, const char* path) {
    if (!m || !path) return fail(SMORE_E_INVALID, "bad argument");
    if (int rc = ensure_device()) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(SMORE_E_IO, "cannot create %s", path);
    CkptHeader h{};
    memcpy(h.magic, kCkptMagic, 8);
    h.version = 2;
    h.dtype = (uint32_t)m->dtype;
    h.n_tables = (uint32_t)m->n_tables;
    h.dim = (uint32_t)m->dim;
    h.V = m->g->V;
    h.rows = m->rows;
    h.rank = m->g->rank;
    h.world = m->g->world;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    const CkptProgress pr{m->ck_seed, m->ck_next_stream, m->ck_sched_total, m->ck_sched_done};
    ok = ok && fwrite(&pr, sizeof(pr), 1, f) == 1;
    const size_t row_bytes = (size_t)m->dim * m->elem();
    const int64_t chunk = std::max<int64_t>(1, (int64_t)((64ull << 20) / row_bytes));
    std::vector<char> host((size_t)chunk * row_bytes);
    for (int t = 0; ok && t < m->n_tables; ++t)
        for (int64_t first = 0; ok && first < m->rows; first += chunk) {
            const int64_t n = std::min(chunk, m->rows - first);
            const int rc = for_row_ranges(m, t, first, n, [&](char* dptr, int64_t f0, int64_t k) -> int {
                CU(cudaMemcpy(host.data() + (size_t)(f0 - first) * row_bytes, dptr, (size_t)k * row_bytes, cudaMemcpyDeviceToHost));
                return SMORE_OK;
            });
            if (rc) {
                fclose(f);
                return rc;
            }
            ok = fwrite(host.data(), row_bytes, (size_t)n, f) == (size_t)n;
        }
    ok = (fclose(f) == 0) && ok;
    return ok ? SMORE_OK : fail(SMORE_E_IO, "short write to %s", path);
}

int smore_model_load_checkpoint(smore_model_t m, const char* path) {
    if (!m || !path) return fail(SMORE_E_INVALID, "bad argument");
    if (int rc = ensure_device()) return rc;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(SMORE_E_IO, "cannot open %s", path);
    CkptHeader h{};
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, kCkptMagic, 8) != 0 || (h.version != 1 && h.version != 2)) {
        fclose(f);
        return fail(SMORE_E_IO, "%s is not a smore_b200 checkpoint", path);
    }
    CkptProgress pr{0, 0, 0, 0};  // a version-1 file holds tables only: loading it is a warm start (progress stays 0 / 0)
    if (h.version == 2 && fread(&pr, sizeof(pr), 1, f) != 1) {
        fclose(f);
        return fail(SMORE_E_IO, "checkpoint %s is truncated", path);
    }
    if ((int)h.dtype != m->dtype || (int)h.n_tables != m->n_tables || (int)h.dim != m->dim || h.V != m->g->V || h.rows != m->rows ||
        h.rank != m->g->rank || h.world != m->g->world) {
        fclose(f);
        return fail(SMORE_E_INVALID, "checkpoint %s does not match the model (V %lld dim %u tables %u dtype %u rank %d/%d)", path,
                    (long long)h.V, h.dim, h.n_tables, h.dtype, h.rank, h.world);
    }
    const size_t row_bytes = (size_t)m->dim * m->elem();
    const int64_t chunk = std::max<int64_t>(1, (int64_t)((64ull << 20) / row_bytes));
    std::vector<char> host((size_t)chunk * row_bytes);
    for (int t = 0; t < m->n_tables; ++t)
        for (int64_t first = 0; first < m->rows; first += chunk) {
            const int64_t n = std::min(chunk, m->rows - first);
            if (fread(host.data(), row_bytes, (size_t)n, f) != (size_t)n) {
                fclose(f);
                return fail(SMORE_E_IO, "checkpoint %s is truncated", path);
            }
            const int rc = for_row_ranges(m, t, first, n, [&](char* dptr, int64_t f0, int64_t k) -> int {
                CU(cudaMemcpy(dptr, host.data() + (size_t)(f0 - first) * row_bytes, (size_t)k * row_bytes, cudaMemcpyHostToDevice));
                return SMORE_OK;
            });
            if (rc) {
                fclose(f);
                return rc;
            }
        }
    fclose(f);
    m->ck_seed = pr.seed;
    m->ck_next_stream = pr.next_stream;
    m->ck_sched_total = pr.sched_total;
    m->ck_sched_done = pr.sched_done;
    return SMORE_OK;
}

int smore_model_progress(smore_model_t m, uint64_t* seed, uint64_t* next_stream, uint64_t* sched_total, uint64_t* sched_done) {
    if (!m) return fail(SMORE_E_INVALID, "null model");
    if (seed) *seed = m->ck_seed;
    if (next_stream) *next_stream = m->ck_next_stream;
    if (sched_total) *sched_total = m->ck_sched_total;
    if (sched_done) *sched_done = m->ck_sched_done;
    return SMORE_OK;
}

int smore_progress(smore_model_t m, uint64_t* done, uint64_t* total, double* alpha, int* running) {
    // no CUDA call, no lock, no thread-local state besides the error string: safe next to a blocking train call on `m`
    if (!m) return fail(SMORE_E_INVALID, "null model");
    unsigned long long* live = __atomic_load_n(&m->live, __ATOMIC_ACQUIRE);
    unsigned long long v[4] = {0, 0, 0, 0};
    if (live) {
        v[3] = __atomic_load_n(&live[3], __ATOMIC_ACQUIRE);
        for (int k = 0; k < 3; ++k) v[k] = __atomic_load_n(&live[k], __ATOMIC_RELAXED);
    }
    if (done) *done = std::min<uint64_t>(v[0], v[2]);
    if (total) *total = v[2];
    if (alpha) memcpy(alpha, &v[1], sizeof(double));
    if (running) *running = (int)v[3];
    return SMORE_OK;
}

int64_t smore_format_rows(const double* rows, int64_t n, int dim, int64_t first_id, int format, char* out, int64_t cap) {
    if (!rows || n < 0 || dim <= 0 || (format != 0 && format != 1)) return fail(SMORE_E_INVALID, "bad argument");
    std::vector<std::string> parts((size_t)host_threads());
    parallel_for(n, 256, [&](int64_t b, int64_t e, int t) {
        format_rows(rows + (size_t)b * (size_t)dim, e - b, dim, nullptr, first_id + b, format, parts[(size_t)t]);
    });
    int64_t total = 0;
    for (auto& s : parts) {
        if (out && total + (int64_t)s.size() <= cap) memcpy(out + total, s.data(), s.size());
        total += (int64_t)s.size();
    }
    return total;
}

// ---- training -----------------------------------------------------------------------------------------------------
int smore_train_line(smore_model_t m, const smore_train_params* p) {
    if (int rc = check_train(m, p, p && p->order == 1 ? 1 : 2, true)) return rc;
    if (p->negative_samples < 0 || p->negative_samples > 31) return fail(SMORE_E_UNSUPPORTED, "negative_samples must be in [0,31]");
    if (p->order != 1 && p->order != 2) return fail(SMORE_E_INVALID, "order must be 1 or 2");
    if (m->xch && m->g->world > 1) {  // bulk-exchange mode, one process per GPU: NCCL carries the row batches
        ExchTransport* tr = exch_nccl_transport(m->g->rank, m->g->world);
        if (!tr) return SMORE_E_INVALID;
        return m->dtype == SMORE_F64 ? train_line_exchange_t<double>(&m, 1, p, *tr) : train_line_exchange_t<float>(&m, 1, p, *tr);
    }
    return m->dtype == SMORE_F64 ? train_line_t<double>(m, p) : train_line_t<float>(m, p);
}

int smore_model_enable_exchange(smore_model_t m, int64_t superbatch, double hot_threshold) {
    if (!m) return fail(SMORE_E_INVALID, "null model");
    smore_graph_s* g = m->g;
    if (g->world == 1) return fail(SMORE_E_INVALID, "the exchange mode only makes sense on a row-sharded graph");
    if (g->rotating || g->synthetic) return fail(SMORE_E_INVALID, "the graph is set up for rotating shards");
    if (superbatch <= 0) superbatch = 1 << 20;
    if (superbatch > (1ll << 27)) return fail(SMORE_E_INVALID, "superbatch must be <= 2^27 samples");
    if (int rc = ensure_device()) return rc;
    if (!m->xch) m->xch = new smore_exchange_s();
    smore_exchange_s* x = m->xch;
    x->superbatch = superbatch;
    x->hot_threshold = hot_threshold;
    x->n_hot = 0;
    if (hot_threshold >= 0) {
        // P(source = v) of the unsharded sampler (C++: out_deg^0.75 / sum, Go: out_deg / sum) x samples of one super-batch
        const double pw = g->sem == SMORE_SEM_CPP ? 0.75 : 1.0;
        const int64_t V = g->V;
        std::vector<double> ps((size_t)V);
        double sum = 0;
        for (int64_t v = 0; v < V; ++v) {
            ps[(size_t)v] = g->out_deg[(size_t)v] > 0 ? std::pow(g->out_deg[(size_t)v], pw) : 0.0;
            sum += ps[(size_t)v];
        }
        const double per_sb = (double)superbatch * (double)g->world;
        std::vector<uint32_t> bits((size_t)(V + 31) / 32, 0u);
        if (sum > 0)
            for (int64_t v = 0; v < V; ++v)
                if (ps[(size_t)v] / sum * per_sb >= hot_threshold && ps[(size_t)v] > 0) {
                    bits[(size_t)(v >> 5)] |= 1u << (v & 31);
                    x->n_hot++;
                }
        if (x->n_hot) {
            if (int rc = x->hot.ensure(bits.size() * 4)) return rc;
            CU(cudaMemcpy(x->hot.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
        }
    }
    return SMORE_OK;
}

int smore_train_line_group(const smore_model_t* shards, int n, const smore_train_params* p) {
    if (!shards || n < 2 || n > kMaxWorld) return fail(SMORE_E_INVALID, "need the 2..8 shards of one model");
    smore_model_s* ms[kMaxWorld];
    for (int r = 0; r < n; ++r) {
        smore_model_s* m = shards[r];
        if (!m) return fail(SMORE_E_INVALID, "null shard");
        if (int rc = check_train(m, p, p && p->order == 1 ? 1 : 2, true)) return rc;
        if (!m->xch) return fail(SMORE_E_INVALID, "shard %d: call smore_model_enable_exchange first", r);
        if (m->g->world != n || m->g->rank != r) return fail(SMORE_E_INVALID, "shards[%d] is rank %d of %d", r, m->g->rank, m->g->world);
        if (m->dim != shards[0]->dim || m->dtype != shards[0]->dtype || m->xch->superbatch != shards[0]->xch->superbatch)
            return fail(SMORE_E_INVALID, "shards disagree on dim / dtype / superbatch");
        ms[r] = m;
    }
    if (p->negative_samples < 0 || p->negative_samples > 31) return fail(SMORE_E_UNSUPPORTED, "negative_samples must be in [0,31]");
    if (p->order != 1 && p->order != 2) return fail(SMORE_E_INVALID, "order must be 1 or 2");
    ExchTransport* tr = exch_local_transport();
    return ms[0]->dtype == SMORE_F64 ? train_line_exchange_t<double>(ms, n, p, *tr) : train_line_exchange_t<float>(ms, n, p, *tr);
}

int smore_exchange_stats(smore_model_t m, uint64_t* superbatches, uint64_t* rows_requested, int64_t* hot_vertices) {
    if (!m || !m->xch) return fail(SMORE_E_INVALID, "exchange mode not enabled on this model");
    if (hot_vertices) *hot_vertices = m->xch->n_hot;
    if (superbatches) *superbatches = m->xch->st_superbatches;
    if (rows_requested) *rows_requested = m->xch->st_rows_moved;
    return SMORE_OK;
}

static int train_walk_common(smore_model_t m, const smore_train_params* p, int walklets) {
    if (int rc = check_train(m, p, 2)) return rc;
    if (p->negative_samples < 0 || p->negative_samples > 31) return fail(SMORE_E_UNSUPPORTED, "negative_samples must be in [0,31]");
    if (p->walk_steps < 0 || p->walk_steps + 1 > kMaxWalkLen) return fail(SMORE_E_UNSUPPORTED, "walk_steps must be < %d", kMaxWalkLen);
    if (p->window_max < 1 || p->window_max > 255) return fail(SMORE_E_INVALID, "window must be in [1,255]");
    if (walklets && (p->window_min < 0 || p->window_min > p->window_max)) return fail(SMORE_E_INVALID, "need 0 <= window_min <= window_max");
    if (walklets && p->semantics != SMORE_SEM_CPP) return fail(SMORE_E_UNSUPPORTED, "Walklets exists only in the C++ tree");
    if (p->walk_times < 1) return fail(SMORE_E_INVALID, "walk_times must be >= 1");
    return m->dtype == SMORE_F64 ? train_walk_t<double>(m, p, walklets) : train_walk_t<float>(m, p, walklets);
}
int smore_train_deepwalk(smore_model_t m, const smore_train_params* p) { return train_walk_common(m, p, 0); }
// Node2Vec.Train (internal/models/node2vec/node2vec.go:176-260): DeepWalk.Train with the biased second-order walk.
int smore_train_node2vec(smore_model_t m, const smore_train_params* p) {
    if (int rc = check_train(m, p, 2)) return rc;
    if (p->semantics != SMORE_SEM_GO) return fail(SMORE_E_UNSUPPORTED, "node2vec exists only in the Go tree");
    if (p->negative_samples < 0 || p->negative_samples > 31) return fail(SMORE_E_UNSUPPORTED, "negative_samples must be in [0,31]");
    if (p->walk_steps < 0 || p->walk_steps + 1 > kMaxWalkLen) return fail(SMORE_E_UNSUPPORTED, "walk_steps must be < %d", kMaxWalkLen);
    if (p->window_max < 1 || p->window_max > 255) return fail(SMORE_E_INVALID, "window must be in [1,255]");
    if (p->walk_times < 1) return fail(SMORE_E_INVALID, "walk_times must be >= 1");
    if (!(p->n2v_p > 0) || !(p->n2v_q > 0)) return fail(SMORE_E_INVALID, "node2vec: p and q must be > 0");
    smore_graph_s* g = m->g;
    if (!g->d_col_sorted) {  // first use: sorted adjacency slices (membership tests) and the raw weights go to the device
        std::vector<int32_t> sorted(g->col);
        parallel_for(g->V, 1 << 12, [&](int64_t vb, int64_t ve, int) {
            for (int64_t v = vb; v < ve; ++v) std::sort(sorted.begin() + g->row_off[(size_t)v], sorted.begin() + g->row_off[(size_t)v + 1]);
        });
        if (int rc = dev_alloc_copy(&g->d_col_sorted, sorted.data(), sorted.size())) return rc;
        if (int rc = dev_alloc_copy(&g->d_w, g->w.data(), g->w.size())) return rc;
    }
    return m->dtype == SMORE_F64 ? train_walk_t<double>(m, p, 0, 1) : train_walk_t<float>(m, p, 0, 1);
}
int smore_train_walklets(smore_model_t m, const smore_train_params* p) { return train_walk_common(m, p, 1); }

int smore_train_hpe(smore_model_t m, const smore_train_params* p) {
    if (int rc = check_train(m, p, 2)) return rc;
    if (p->semantics != SMORE_SEM_CPP) return fail(SMORE_E_UNSUPPORTED, "HPE with walks exists only in the C++ tree (the Go hpe model is LINE-2: use smore_train_line)");
    if (p->negative_samples < 0 || p->negative_samples > 31) return fail(SMORE_E_UNSUPPORTED, "negative_samples must be in [0,31]");
    if (p->walk_steps < 1) return fail(SMORE_E_INVALID, "walk_steps must be >= 1");
    return m->dtype == SMORE_F64 ? train_hpe_t<double>(m, p) : train_hpe_t<float>(m, p);
}

int smore_train_mf(smore_model_t m, const smore_train_params* p) {
    if (int rc = check_train(m, p, 1)) return rc;
    if (p->semantics != SMORE_SEM_CPP) return fail(SMORE_E_UNSUPPORTED, "MF exists only in the C++ tree");
    if (p->negative_samples < 0 || p->negative_samples > 31) return fail(SMORE_E_UNSUPPORTED, "negative_samples must be in [0,31]");
    return m->dtype == SMORE_F64 ? train_mf_t<double>(m, p) : train_mf_t<float>(m, p);
}

int smore_train_bpr(smore_model_t m, const smore_train_params* p) {
    if (int rc = check_train(m, p, p && p->semantics == SMORE_SEM_GO ? 2 : 1)) return rc;
    return m->dtype == SMORE_F64 ? train_ranking_t<double>(m, p, RANK_BPR) : train_ranking_t<float>(m, p, RANK_BPR);
}
int smore_train_skewopt(smore_model_t m, const smore_train_params* p) {
    if (int rc = check_train(m, p, 1)) return rc;
    if (p->semantics != SMORE_SEM_CPP) return fail(SMORE_E_UNSUPPORTED, "Skew-OPT (SPR) exists only in the C++ tree");
    if (p->eta < 1 || p->eta > 15) return fail(SMORE_E_INVALID, "eta must be in [1,15]");
    if (!(p->omega != 0)) return fail(SMORE_E_INVALID, "omega must be non-zero");
    return m->dtype == SMORE_F64 ? train_ranking_t<double>(m, p, RANK_SKEWOPT) : train_ranking_t<float>(m, p, RANK_SKEWOPT);
}

// ---- CPR / TPR (Go tree only): a second graph and a third table ------------------------------------------------------
int smore_model_attach_aux(smore_model_t m, int64_t V_aux, const int64_t* row_off, const int32_t* col, const double* rows,
                           uint64_t seed) {
    if (!m || V_aux < 1 || !row_off) return fail(SMORE_E_INVALID, "bad model / auxiliary graph");
    if (m->g->world != 1) return fail(SMORE_E_UNSUPPORTED, "CPR / TPR run on unsharded graphs");
    const int64_t E = row_off[V_aux];
    if (row_off[0] != 0 || E < 0 || (E > 0 && !col)) return fail(SMORE_E_INVALID, "bad auxiliary CSR");
    for (int64_t v = 0; v < V_aux; ++v)
        if (row_off[v + 1] < row_off[v]) return fail(SMORE_E_INVALID, "auxiliary row_off must be non-decreasing");
    for (int64_t e = 0; e < E; ++e)
        if (col[e] < 0 || col[e] >= V_aux) return fail(SMORE_E_INVALID, "auxiliary col[%lld] = %d is not a vertex", (long long)e, col[e]);
    if (int rc = ensure_device()) return rc;
    cudaFree(m->d_aux_off); cudaFree(m->d_aux_col); cudaFree(m->aux_tab);
    m->d_aux_off = nullptr; m->d_aux_col = nullptr; m->aux_tab = nullptr;
    m->aux_V = V_aux;
    m->aux_E = E;
    if (int rc = dev_alloc_copy(&m->d_aux_off, row_off, (size_t)V_aux + 1)) return rc;
    const int32_t none = 0;
    if (int rc = dev_alloc_copy(&m->d_aux_col, E ? col : &none, (size_t)std::max<int64_t>(E, 1))) return rc;
    const size_t cnt = (size_t)V_aux * (size_t)m->dim;
    CU(cudaMalloc(&m->aux_tab, cnt * m->elem()));
    if (rows) {
        if (m->dtype == SMORE_F64) {
            CU(cudaMemcpy(m->aux_tab, rows, cnt * sizeof(double), cudaMemcpyHostToDevice));
        } else {
            std::vector<float> f(cnt);
            parallel_for((int64_t)cnt, 1 << 16, [&](int64_t b, int64_t e, int) {
                for (int64_t i = b; i < e; ++i) f[(size_t)i] = (float)rows[i];
            });
            CU(cudaMemcpy(m->aux_tab, f.data(), cnt * sizeof(float), cudaMemcpyHostToDevice));
        }
    } else {  // (U - 0.5) / dim like every Go table (cpr.go:118-124, tpr.go:92-98), the init stream of a third table
        const int blocks = (int)std::min<size_t>((cnt + 255) / 256, 148 * 8);
        if (m->dtype == SMORE_F64) k_init_table<double><<<blocks, 256>>>((double*)m->aux_tab, (int64_t)cnt, m->dim, seed, kInitStreamBase + 2, 1, 0, 0, 0);
        else k_init_table<float><<<blocks, 256>>>((float*)m->aux_tab, (int64_t)cnt, m->dim, seed, kInitStreamBase + 2, 1, 0, 0, 0);
        g_launches++;
        CU(cudaGetLastError());
        CU(cudaDeviceSynchronize());
    }
    return SMORE_OK;
}

int smore_model_get_aux_rows(smore_model_t m, int64_t first, int64_t n, double* host) {
    if (!m || !m->aux_tab || !host) return fail(SMORE_E_INVALID, "no auxiliary table attached / null buffer");
    if (first < 0 || n < 0 || first + n > m->aux_V) return fail(SMORE_E_INVALID, "row range out of bounds (the auxiliary table holds %lld rows)", (long long)m->aux_V);
    const size_t cnt = (size_t)n * (size_t)m->dim;
    if (cnt == 0) return SMORE_OK;
    if (m->dtype == SMORE_F64) {
        CU(cudaMemcpy(host, (const double*)m->aux_tab + (size_t)first * m->dim, cnt * sizeof(double), cudaMemcpyDeviceToHost));
    } else {
        std::vector<float> f(cnt);
        CU(cudaMemcpy(f.data(), (const float*)m->aux_tab + (size_t)first * m->dim, cnt * sizeof(float), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < cnt; ++i) host[i] = (double)f[i];
    }
    return SMORE_OK;
}

static int train_aux_common(smore_model_t m, const smore_train_params* p, int kind, const char* name) {
    if (int rc = check_train(m, p, 2)) return rc;
    if (p->semantics != SMORE_SEM_GO) return fail(SMORE_E_UNSUPPORTED, "%s exists only in the Go tree", name);
    if (!m->aux_tab) return fail(SMORE_E_INVALID, "%s: call smore_model_attach_aux first (second graph + third table)", name);
    return m->dtype == SMORE_F64 ? train_ranking_t<double>(m, p, kind) : train_ranking_t<float>(m, p, kind);
}
int smore_train_cpr(smore_model_t m, const smore_train_params* p) { return train_aux_common(m, p, RANK_CPR, "CPR"); }
int smore_train_tpr(smore_model_t m, const smore_train_params* p) {
    if (p && !(p->text_weight >= 0 && p->text_weight <= 1)) return fail(SMORE_E_INVALID, "text_weight must be in [0,1]");
    return train_aux_common(m, p, RANK_TPR, "TPR");
}

int smore_train_warp(smore_model_t m, const smore_train_params* p) {
    if (int rc = check_train(m, p, 1)) return rc;
    if (p->semantics != SMORE_SEM_CPP) return fail(SMORE_E_UNSUPPORTED, "WARP exists only in the C++ tree");
    return m->dtype == SMORE_F64 ? train_ranking_t<double>(m, p, RANK_WARP) : train_ranking_t<float>(m, p, RANK_WARP);
}
int smore_train_hoprec(smore_model_t m, const smore_train_params* p) {
    if (int rc = check_train(m, p, 1)) return rc;
    if (p->semantics != SMORE_SEM_CPP) return fail(SMORE_E_UNSUPPORTED, "HOP-Rec exists only in the C++ tree");
    if (!m->g->has_field) return fail(SMORE_E_INVALID, "HOP-Rec needs field data (smore_graph_load_field / set_field)");
    if (p->walk_steps < 1) return fail(SMORE_E_INVALID, "walk_steps must be >= 1");
    return m->dtype == SMORE_F64 ? train_ranking_t<double>(m, p, RANK_HOPREC) : train_ranking_t<float>(m, p, RANK_HOPREC);
}

int smore_train_stats(smore_model_t m, uint64_t* samples, uint64_t* pair_updates, uint64_t* words_stream0,
                      double* warp_mean_tries, double* kernel_ms) {
    if (!m) return fail(SMORE_E_INVALID, "null model");
    if (samples) *samples = m->st_samples;
    if (pair_updates) *pair_updates = m->st_pairs;
    if (words_stream0) *words_stream0 = m->st_words0;
    if (warp_mean_tries) *warp_mean_tries = m->st_samples ? (double)m->st_tries / (double)m->st_samples : 0.0;
    if (kernel_ms) *kernel_ms = m->st_ms;
    return SMORE_OK;
}

}  // extern "C"
