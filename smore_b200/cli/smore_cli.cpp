// Host-side mirror of the reference CLIs on top of the C ABI (include/smore_b200.h).
//
//   smore <model> -train net.txt -save rep.txt [flags]      model in {line, deepwalk, walklets, bpr, warp, hoprec, hpe, mf, skewopt}
//   (or invoke through a symlink named after the model: `line -train ...`)
//
// Flag names, defaults and the four-call sequence LoadEdgeList -> Init -> Train -> SaveWeights are those of
// cmd/{line,deepwalk,bpr}/main.go (Go tree) and cli/{line,deepwalk,walklets,bpr,warp,hoprec}.cpp (C++ tree). Both flag
// spellings are accepted (`-undirected 1` as the C++ ArgPos parser takes it, `-undirected=false` / bare `-undirected`
// as Go's flag package does). Extra flags: -semantics {go,cpp} (which tree's maths; default go where a Go CLI exists),
// -mode {hogwild,deterministic}, -seed N, -dtype {f32,f64}. -threads is accepted and ignored: the workers are GPU warps.
// The Go toolchain is not available in the build image; INTEGRATION.md has the cgo binding that calls the same ABI.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/smore_b200.h"

namespace {

struct Args {
    std::map<std::string, std::string> kv;
    bool has(const std::string& k) const { return kv.count(k) != 0; }
    std::string str(const std::string& k, const std::string& d) const { return has(k) ? kv.at(k) : d; }
    long num(const std::string& k, long d) const { return has(k) ? atol(kv.at(k).c_str()) : d; }
    double real(const std::string& k, double d) const { return has(k) ? atof(kv.at(k).c_str()) : d; }
    bool flag(const std::string& k, bool d) const {
        if (!has(k)) return d;
        const std::string& v = kv.at(k);
        return !(v == "0" || v == "false" || v == "False" || v == "FALSE" || v == "f");
    }
};

bool looks_bool(const char* s) {
    static const char* vals[] = {"0", "1", "true", "false", "True", "False", "TRUE", "FALSE", "t", "f"};
    for (const char* v : vals)
        if (!strcmp(s, v)) return true;
    return false;
}

Args parse(int argc, char** argv, int first) {
    Args a;
    for (int i = first; i < argc; ++i) {
        const char* s = argv[i];
        if (s[0] != '-') continue;
        while (*s == '-') ++s;
        std::string key(s), val;
        size_t eq = key.find('=');
        if (eq != std::string::npos) {
            val = key.substr(eq + 1);
            key = key.substr(0, eq);
        } else if (key == "undirected") {  // Go bool flag: bare means true; C++: `-undirected 1`
            if (i + 1 < argc && looks_bool(argv[i + 1])) val = argv[++i];
            else val = "true";
        } else if (i + 1 < argc) {
            val = argv[++i];
        }
        a.kv[key] = val;
    }
    return a;
}

int die(const char* what) {
    fprintf(stderr, "smore: %s: %s\n", what, smore_last_error());
    return 1;
}

void usage() {
    printf("[smore_b200]\n\tB200-native SMORe trainers (LINE, DeepWalk, Walklets, BPR, WARP, HOP-Rec)\n\n"
           "Usage:\n\tsmore <line|deepwalk|walklets|node2vec|bpr|warp|hoprec|hpe|mf|skewopt> -train net.txt -save rep.txt [options]\n"
           "\tsmore cpr -train_target t.txt -train_source s.txt -save_user u.txt -save_target t.rep -save_source s.rep\n"
           "\t          [-update_times 10 -alpha 0.1 -user_reg 0.01 -item_reg 0.01 -margin 8]\n"
           "\tsmore tpr -train_ui ui.txt -train_iw iw.txt -save_user u.txt -save_item i.txt -save_word w.txt\n"
           "\t          [-sample_times 10 -alpha 0.025 -lambda 0.025 -text_weight 0.5]\n\n"
           "Options Description:\n"
           "\t-train <string>\n\t\tTrain the Network data\n"
           "\t-save <string>\n\t\tSave the representation data\n"
           "\t-field <string>\n\t\tField data (hoprec)\n"
           "\t-dimensions <int>\n\t\tDimension of vertex representation; default is 64\n"
           "\t-undirected <bool>\n\t\tWhether the edge is undirected; default is 1 (line, deepwalk, walklets), 0 (bpr, warp)\n"
           "\t-negative_samples <int>\n\t\tNumber of negative examples; default is 5\n"
           "\t-order <int>\n\t\tLINE: order of proximity (1 or 2); default is 2\n"
           "\t-sample_times <int>\n\t\tNumber of training samples (cpp: *Million, go: *edge lines); default is 10\n"
           "\t-walk_times <int> -walk_steps <int> -window_size <int> -window_min <int> -window_max <int>\n"
           "\t-p <float> -q <float>\n\t\tnode2vec return / in-out parameters; defaults 1, 1 (window 10, walk_steps 80)\n"
           "\t-lambda <float>\n\t\tGo BPR regularisation; default is 0.001\n"
           "\t-reg <float>\n\t\tHPE / MF regularisation; default is 0.01\n"
           "\t-xi <float> -omega <float> -eta <int>\n\t\tSkew-OPT margin shift / scale / power; defaults 10, 3, 3\n"
           "\t-alpha <float>\n\t\tInit learning rate; default is 0.025\n"
           "\t-threads <int>\n\t\tAccepted for compatibility; workers are GPU warps\n"
           "\t-load_v <string> -load_c <string>\n\t\tPretrained vertex / context embeddings (text, matched by vertex name)\n"
           "\t-checkpoint <string> -resume <string>\n\t\tWrite / read a binary snapshot of all tables\n"
           "\t-semantics <go|cpp> -mode <hogwild|deterministic> -seed <int> -dtype <f32|f64> -device <int>\n");
}


// CPR / TPR (Go tree only; cmd/cpr/main.go, cmd/tpr/main.go): two edge lists, three output files.
int write_named_rows(const std::string& path, smore_graph_t names, const std::vector<double>& rows, int64_t n, int dim) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) {
        fprintf(stderr, "smore: cannot write %s\n", path.c_str());
        return 1;
    }
    fprintf(f, "%lld %d\n", (long long)n, dim);
    for (int64_t v = 0; v < n; ++v) {
        const char* nm = smore_graph_vertex_name(names, v);
        fputs(nm ? nm : "", f);
        for (int d = 0; d < dim; ++d) fprintf(f, " %.6f", rows[(size_t)v * dim + d]);
        fputc('\n', f);
    }
    fclose(f);
    return 0;
}

int run_two_graph(const std::string& model, Args& a) {
    const bool cpr = model == "cpr";
    const std::string f1 = a.str(cpr ? "train_target" : "train_ui", ""), f2 = a.str(cpr ? "train_source" : "train_iw", "");
    const std::string s0 = a.str("save_user", ""), s1 = a.str(cpr ? "save_target" : "save_item", ""),
                      s2 = a.str(cpr ? "save_source" : "save_word", "");
    if (f1.empty() || f2.empty()) {
        fprintf(stderr, cpr ? "smore: cpr needs -train_target and -train_source\n" : "smore: tpr needs -train_ui and -train_iw\n");
        return 1;
    }
    if (a.str("semantics", "go") != "go") {
        fprintf(stderr, "smore: %s exists only in the Go tree; use -semantics go\n", model.c_str());
        return 1;
    }
    const int dim = (int)a.num("dimensions", 64);
    const bool undirected = a.flag("undirected", true);
    const int dtype = a.str("dtype", "f32") == "f64" ? SMORE_F64 : SMORE_F32;
    if (smore_init((int)a.num("device", 0))) return die("init");
    smore_graph_t g1 = nullptr, g2 = nullptr;
    printf(cpr ? "Loading target domain...\n" : "Loading user-item graph...\n");
    if (smore_graph_load_edge_list(f1.c_str(), undirected, SMORE_SEM_GO, SMORE_NEG_DEGREES, &g1)) return die("LoadEdgeList");
    printf(cpr ? "Loading source domain...\n" : "Loading item-word graph...\n");
    if (smore_graph_load_edge_list(f2.c_str(), undirected, SMORE_SEM_GO, SMORE_NEG_DEGREES, &g2)) return die("LoadEdgeList");
    int64_t V1 = 0, E1 = 0, L1 = 0, V2 = 0, E2 = 0, L2 = 0;
    smore_graph_info(g1, &V1, &E1, &L1);
    smore_graph_info(g2, &V2, &E2, &L2);
    std::vector<int64_t> off2((size_t)V2 + 1);
    std::vector<int32_t> col2((size_t)std::max<int64_t>(E2, 1));
    if (smore_graph_get_csr(g2, off2.data(), col2.data(), nullptr)) return die("second graph");

    smore_train_params p;
    smore_train_params_default(&p);
    p.semantics = SMORE_SEM_GO;
    p.mode = a.str("mode", "hogwild") == "deterministic" ? SMORE_MODE_DETERMINISTIC : SMORE_MODE_HOGWILD;
    p.seed = (uint64_t)a.num("seed", 1);
    if (cpr) {  // cmd/cpr/main.go:19-24
        p.total = (uint64_t)a.num("update_times", 10) * 1000000ull;
        p.alpha = a.real("alpha", 0.1);
        p.lambda = a.real("user_reg", 0.01);
        p.item_reg = a.real("item_reg", 0.01);
        p.margin = a.real("margin", 8.0);
    } else {  // cmd/tpr/main.go:19-22
        p.total = (uint64_t)a.num("sample_times", 10) * (uint64_t)L1;
        p.alpha = a.real("alpha", 0.025);
        p.lambda = a.real("lambda", 0.025);
        p.text_weight = a.real("text_weight", 0.5);
    }
    smore_model_t m = nullptr;
    if (smore_model_create(g1, dim, 2, dtype, &m)) return die("Init");
    if (smore_model_init(m, 0, 1, p.seed) || smore_model_init(m, 1, 1, p.seed)) return die("Init");
    if (smore_model_attach_aux(m, V2, off2.data(), col2.data(), nullptr, p.seed)) return die("Init");
    printf("Model Setting:\n\tdimension:\t\t%d\n", dim);
    if (cpr) {
        printf("\tmax_user_id:\t\t%lld\n\ttarget_items:\t\t%lld\n\tsource_items:\t\t%lld\n", (long long)std::max(V1, V2), (long long)V1, (long long)V2);
        if (V2 > V1)
            printf("\t(the reference also keeps %lld user rows past the target graph: never trained, unnamed, not written here)\n",
                   (long long)(V2 - V1));
        printf("Model:\n\t[CPR: Cross-Domain Preference Ranking]\nLearning Parameters:\n\tupdate_times:\t\t%ld\n\talpha:\t\t\t%.6f\n"
               "\tuser_reg:\t\t%.6f\n\titem_reg:\t\t%.6f\n\tmargin:\t\t\t%.6f\n", a.num("update_times", 10), p.alpha, p.lambda, p.item_reg, p.margin);
    } else {
        printf("\ttext_weight:\t\t%.2f\n\tnum_users:\t\t%lld\n\tnum_items:\t\t%lld\n\tnum_words:\t\t%lld\n", p.text_weight, (long long)V1,
               (long long)V1, (long long)V2);
        printf("Model:\n\t[TPR: Text-aware Preference Ranking]\nLearning Parameters:\n\tsample_times:\t\t%ld\n\talpha:\t\t\t%.6f\n"
               "\tlambda:\t\t\t%.6f\n", a.num("sample_times", 10), p.alpha, p.lambda);
    }
    if (a.has("threads")) printf("\tworkers:\t\t-threads %ld ignored: workers are GPU warps\n", a.num("threads", 1));
    printf("Start Training:\n");
    std::atomic<bool> stop{false};
    std::thread poll([&] {
        while (!stop.load(std::memory_order_acquire)) {
            uint64_t done = 0, total = 0;
            double alpha = 0;
            int running = 0;
            if (smore_progress(m, &done, &total, &alpha, &running) == 0 && running && total) {
                printf("\tAlpha: %.6f\tProgress: %.3f %%%c", alpha, 100.0 * (double)done / (double)total, 13);
                fflush(stdout);
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
    });
    const int rc = cpr ? smore_train_cpr(m, &p) : smore_train_tpr(m, &p);
    stop.store(true, std::memory_order_release);
    poll.join();
    if (rc) return die("Train");
    uint64_t samples = 0, pairs = 0;
    double ms = 0, alpha_end = 0;
    smore_train_stats(m, &samples, &pairs, nullptr, nullptr, &ms);
    smore_progress(m, nullptr, nullptr, &alpha_end, nullptr);
    printf("\tAlpha: %.6f\tProgress: 100.00 %%\n", alpha_end);
    printf("\t%llu samples, %llu updates in %.1f ms on the device (%.1f M samples/s)\n", (unsigned long long)samples,
           (unsigned long long)pairs, ms, ms > 0 ? samples / ms / 1e3 : 0.0);
    printf("Save Model:\n");
    if (!s0.empty()) {
        if (smore_model_save_weights(m, 0, s0.c_str(), 1)) return die("SaveWeights");
        printf("\tSave users to <%s>\n", s0.c_str());
    }
    if (!s1.empty()) {
        if (smore_model_save_weights(m, 1, s1.c_str(), 1)) return die("SaveWeights");
        printf("\tSave %s to <%s>\n", cpr ? "target items" : "items", s1.c_str());
    }
    if (!s2.empty()) {
        std::vector<double> rows((size_t)V2 * (size_t)dim);
        if (smore_model_get_aux_rows(m, 0, V2, rows.data())) return die("SaveWeights");
        if (write_named_rows(s2, g2, rows, V2, dim)) return 1;
        printf("\tSave %s to <%s>\n", cpr ? "source items" : "words", s2.c_str());
    }
    smore_model_destroy(m);
    smore_graph_destroy(g1);
    smore_graph_destroy(g2);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    std::string model;
    int first = 1;
    const char* base = strrchr(argv[0], '/');
    base = base ? base + 1 : argv[0];
    for (const char* m : {"line", "deepwalk", "walklets", "node2vec", "bpr", "warp", "hoprec", "hpe", "mf", "skewopt", "cpr", "tpr"})
        if (!strcmp(base, m)) model = m;
    if (model.empty()) {
        if (argc < 2 || argv[1][0] == '-') {
            usage();
            return argc < 2 ? 0 : 1;
        }
        model = argv[1];
        first = 2;
    }
    if (argc <= first) {
        usage();
        return 0;
    }
    Args a = parse(argc, argv, first);
    if (model == "cpr" || model == "tpr") return run_two_graph(model, a);
    const bool has_go_cli = model == "line" || model == "deepwalk" || model == "bpr" || model == "node2vec";
    if (!has_go_cli && model != "walklets" && model != "warp" && model != "hoprec" && model != "hpe" && model != "mf" && model != "skewopt") {
        fprintf(stderr, "smore: unknown model '%s'\n", model.c_str());
        return 1;
    }
    const std::string sem_s = a.str("semantics", has_go_cli ? "go" : "cpp");
    const int sem = sem_s == "go" ? SMORE_SEM_GO : SMORE_SEM_CPP;
    if (sem == SMORE_SEM_CPP && model == "node2vec") {
        fprintf(stderr, "smore: node2vec exists only in the Go tree; use -semantics go\n");
        return 1;
    }
    if (sem == SMORE_SEM_GO && !has_go_cli) {
        fprintf(stderr, "smore: %s exists only in the C++ tree; use -semantics cpp\n", model.c_str());
        return 1;
    }
    const std::string train = a.str("train", ""), save = a.str("save", "");
    if (train.empty() || save.empty()) {
        usage();
        return 1;
    }
    // (MF: one table, "no_degrees" negatives and a directed load, like the ranking models: MF.cpp:4-7, cli/mf.cpp:63)
    const bool ranking = model == "bpr" || model == "warp" || model == "hoprec" || model == "mf" || model == "skewopt";
    // defaults: cmd/line/main.go:15-21, cmd/deepwalk/main.go:13-22, cmd/bpr/main.go:13-20, cli/*.cpp
    const int dim = (int)a.num("dimensions", 64);
    const bool undirected = model == "hoprec" ? true : a.flag("undirected", !(model == "bpr" || model == "warp" || model == "mf" || model == "skewopt"));
    const int sample_times = (int)a.num("sample_times", 10);
    const int dtype = a.str("dtype", "f32") == "f64" ? SMORE_F64 : SMORE_F32;

    if (smore_init((int)a.num("device", 0))) return die("init");
    const int neg_method = (sem == SMORE_SEM_CPP && ranking) ? SMORE_NEG_NO_DEGREES : SMORE_NEG_DEGREES;  // BPR.cpp:4-7
    smore_graph_t g = nullptr;
    if (smore_graph_load_edge_list(train.c_str(), undirected, sem, neg_method, &g)) return die("LoadEdgeList");
    int64_t V = 0, E = 0, n_lines = 0;
    smore_graph_info(g, &V, &E, &n_lines);
    printf("Graph loaded: %lld vertices, %lld edge lines, %lld adjacency entries\n", (long long)V,
           (long long)(sem == SMORE_SEM_GO ? n_lines : (undirected ? n_lines / 2 : n_lines)), (long long)E);
    if (model == "hoprec") {
        if (!a.has("field")) {
            fprintf(stderr, "smore: hoprec needs -field\n");
            return 1;
        }
        if (smore_graph_load_field(g, a.str("field", "").c_str())) return die("LoadFieldMeta");
    }

    smore_train_params p;
    smore_train_params_default(&p);
    p.semantics = sem;
    p.mode = a.str("mode", "hogwild") == "deterministic" ? SMORE_MODE_DETERMINISTIC : SMORE_MODE_HOGWILD;
    p.seed = (uint64_t)a.num("seed", 1);
    p.alpha = a.real("alpha", 0.025);
    p.negative_samples = (int)a.num("negative_samples", 5);
    p.order = a.num("order", 2) == 1 ? 1 : 2;
    p.lambda = (model == "hpe" || model == "mf") ? a.real("reg", 0.01) : a.real("lambda", 0.001);  // cli/hpe.cpp:57
    p.xi = a.real("xi", 10.0);  // cli/skewopt.cpp:53-54
    p.omega = a.real("omega", 3.0);
    p.eta = (int)a.num("eta", 3);
    p.walk_times = (int)a.num("walk_times", 10);
    // cli/deepwalk.cpp:56 defaults walk_steps to 5 (sic), cli/walklets.cpp:54 and cmd/deepwalk/main.go:20 to 40
    // cmd/node2vec/main.go:18-22: window_size 10, walk_steps 80, p = q = 1
    p.walk_steps = (int)a.num("walk_steps", model == "node2vec" ? 80 : model == "hoprec" || model == "hpe" ? 5 : (model == "deepwalk" && sem == SMORE_SEM_CPP ? 5 : 40));
    p.window_min = (int)a.num("window_min", model == "walklets" ? 2 : 1);
    p.window_max = (int)a.num(model == "walklets" ? "window_max" : "window_size", model == "node2vec" ? 10 : 5);
    p.n2v_p = a.real("p", 1.0);
    p.n2v_q = a.real("q", 1.0);
    // LINE.cpp:119 (sample_times * 1e6) vs line.go:85 (sample_times * MaxLine)
    p.total = sem == SMORE_SEM_CPP ? (uint64_t)sample_times * 1000000ull : (uint64_t)sample_times * (uint64_t)n_lines;

    // tables: Go models own wVertex + wContext, both random (line.go:52-69); C++ LINE-2: context zeros (LINE.cpp:92),
    // LINE-1 / BPR / WARP / HOP-Rec: one table (LINE.cpp:73-84, BPR.cpp:44-51), DeepWalk: both random (DeepWalk.cpp:44-56)
    int n_tables = 2;
    if (sem == SMORE_SEM_CPP && (ranking || (model == "line" && p.order == 1))) n_tables = 1;
    if (sem == SMORE_SEM_GO && model == "line" && p.order == 1) n_tables = 1;
    smore_model_t m = nullptr;
    if (smore_model_create(g, dim, n_tables, dtype, &m)) return die("Init");
    printf("Model Setting:\n\tdimension:\t\t%d\n", dim);
    if (smore_model_init(m, 0, 1, p.seed)) return die("Init");
    if (n_tables == 2) {
        const bool ctx_random = !(sem == SMORE_SEM_CPP && model == "line");
        if (smore_model_init(m, 1, ctx_random ? 1 : 0, p.seed)) return die("Init");
    }

    if (model == "skewopt") {  // SPR::Init adds 0.01 to every element (src/model/SkewOPT.cpp:49)
        std::vector<double> rows((size_t)V * (size_t)dim);
        if (smore_model_get_rows(m, 0, 0, V, rows.data())) return die("Init");
        for (double& x : rows) x += 0.01;
        if (smore_model_set_rows(m, 0, 0, V, rows.data())) return die("Init");
    }
    // cli/deepwalk.cpp:61-62, DeepWalk.cpp:83-92: pretrained vertex / context embeddings replace the random init
    for (int t = 0; t < n_tables; ++t) {
        const char* key = t == 0 ? "load_v" : "load_c";
        if (!a.has(key)) continue;
        int64_t n = 0;
        printf("\tload %s pretrain:\t%s\n", t == 0 ? "vertex" : "context", a.str(key, "").c_str());
        if (smore_model_load_pretrain(m, t, a.str(key, "").c_str(), &n)) return die("LoadPreTrain");
        printf("\t# of Pre-train:\t\t%lld\n", (long long)n);
    }
    // -resume: continue the interrupted run where its checkpoint left it (same seed, unused sampler streams, same LR
    // schedule); a checkpoint without a position (format version 1) or of another schedule is a warm start
    uint64_t resume_done = 0, resume_stream = 0;
    if (a.has("resume")) {
        if (smore_model_load_checkpoint(m, a.str("resume", "").c_str())) return die("LoadCheckpoint");
        uint64_t ck_seed = 0, ck_stream = 0, ck_total = 0, ck_done = 0;
        smore_model_progress(m, &ck_seed, &ck_stream, &ck_total, &ck_done);
        const bool walk = model == "deepwalk" || model == "walklets" || model == "node2vec";
        if (!walk && ck_total == p.total && ck_done > 0 && ck_done < ck_total) {
            resume_done = ck_done;
            resume_stream = ck_stream;
            if (!a.has("seed")) p.seed = ck_seed;
            printf("\tresume:\t\t\t%llu of %llu samples done, continuing on stream %llu\n", (unsigned long long)ck_done,
                   (unsigned long long)ck_total, (unsigned long long)ck_stream);
        } else {
            printf("\tresume:\t\t\ttables only (checkpoint holds %llu of %llu, this run is %llu): warm start\n",
                   (unsigned long long)ck_done, (unsigned long long)ck_total, (unsigned long long)p.total);
        }
    }

    printf("Model:\n\t[%s] (%s semantics, %s, %s)\n", model.c_str(), sem_s.c_str(),
           p.mode == SMORE_MODE_HOGWILD ? "hogwild" : "deterministic", dtype == SMORE_F64 ? "f64" : "f32");
    printf("Learning Parameters:\n\tsample_times:\t\t%d\n\tnegative_samples:\t%d\n\talpha:\t\t\t%g\n", sample_times,
           p.negative_samples, p.alpha);
    if (a.has("threads")) printf("\tworkers:\t\t-threads %ld ignored: workers are GPU warps\n", a.num("threads", 1));
    printf("Start Training:\n");
    auto train_once = [&](const smore_train_params& q) -> int {
        if (model == "line") return smore_train_line(m, &q);
        if (model == "deepwalk") return smore_train_deepwalk(m, &q);
        if (model == "walklets") return smore_train_walklets(m, &q);
        if (model == "node2vec") return smore_train_node2vec(m, &q);
        if (model == "bpr") return smore_train_bpr(m, &q);
        if (model == "warp") return smore_train_warp(m, &q);
        if (model == "hpe") return smore_train_hpe(m, &q);
        if (model == "mf") return smore_train_mf(m, &q);
        if (model == "skewopt") return smore_train_skewopt(m, &q);
        return smore_train_hoprec(m, &q);
    };
    uint64_t samples = 0, pairs = 0;
    double ms = 0;
    const bool walk_model = model == "deepwalk" || model == "walklets" || model == "node2vec";
    const int chunks = (p.mode == SMORE_MODE_HOGWILD && !walk_model && p.total >= 2000000) ? 20 : 1;
    const long ckpt_every = a.num("checkpoint_every", 0);  // chunks between intermediate checkpoints (0: only at the end)
    double last_alpha = p.alpha;
    const int first_chunk = chunks > 1 ? (int)(resume_done / (p.total / chunks)) : 0;
    // -max_chunks N: stop after N chunks of this invocation (time-sliced jobs: continue later with -resume <checkpoint>)
    const int last_chunk = a.has("max_chunks") ? std::min<long>(chunks, first_chunk + a.num("max_chunks", chunks)) : chunks;
    for (int c = first_chunk; c < last_chunk; ++c) {
        // Hogwild: the run is cut into chunks of one LR schedule so that the reference's progress line can be printed
        // (LINE.cpp:185); the deterministic mode stays one call (one worker, the reference's exact loop).
        smore_train_params q = p;
        if (chunks > 1) {
            q.total = p.total / chunks;
            q.sched_total = p.total;
            q.sched_offset = (uint64_t)c * q.total;
            q.stream_base = std::max<uint64_t>(p.stream_base + (uint64_t)c * (1ull << 24), resume_stream);
        }
        // the reference's workers print "Alpha / Progress" every MONITOR samples (LINE.cpp:179-187); here a host thread
        // polls smore_progress -- the one call that is safe next to the blocking train call -- and prints the same line
        std::atomic<bool> stop{false};
        std::thread poll([&] {
            while (!stop.load(std::memory_order_acquire)) {
                uint64_t done = 0, total = 0;
                double alpha = 0;
                int running = 0;
                if (smore_progress(m, &done, &total, &alpha, &running) == 0 && running && total) {
                    printf("\tAlpha: %.6f\tProgress: %.3f %%%c", alpha, 100.0 * (double)done / (double)total, 13);
                    fflush(stdout);
                }
                std::this_thread::sleep_for(std::chrono::milliseconds(100));
            }
        });
        const int train_rc = train_once(q);
        stop.store(true, std::memory_order_release);
        poll.join();
        if (train_rc) return die("Train");
        if (ckpt_every > 0 && a.has("checkpoint") && (c + 1) % ckpt_every == 0 && c + 1 < chunks &&
            smore_model_save_checkpoint(m, a.str("checkpoint", "").c_str()))
            return die("SaveCheckpoint");
        uint64_t s1 = 0, p1 = 0;
        double ms1 = 0;
        smore_train_stats(m, &s1, &p1, nullptr, nullptr, &ms1);
        samples += s1;
        pairs += p1;
        ms += ms1;
        const double done = (double)(c + 1) / chunks;
        last_alpha = p.alpha * (1.0 - done);
        if (last_alpha < p.alpha * 0.0001) last_alpha = p.alpha * 0.0001;
        if (chunks > 1) {
            printf("\tAlpha: %.6f\tProgress: %.3f %%%c", last_alpha, done * 100, 13);
            fflush(stdout);
        }
    }
    if (last_chunk == chunks) printf("\tAlpha: %.6f\tProgress: 100.00 %%\n", last_alpha);
    else printf("\tAlpha: %.6f\tProgress: %.2f %% (stopped by -max_chunks)\n", last_alpha, 100.0 * last_chunk / chunks);
    printf("\t%llu samples, %llu pair updates in %.1f ms on the device (%.1f M updates/s)\n",
           (unsigned long long)samples, (unsigned long long)pairs, ms, ms > 0 ? pairs / ms / 1e3 : 0.0);

    printf("Save Model:\n");
    // number format: C++ iostream default (%g) vs Go "%.6f" (LINE.cpp:37, line.go:226)
    if (smore_model_save_weights(m, 0, save.c_str(), sem == SMORE_SEM_GO ? 1 : 0)) return die("SaveWeights");
    printf("\tSave to <%s>\n", save.c_str());
    if (a.has("checkpoint") && smore_model_save_checkpoint(m, a.str("checkpoint", "").c_str())) return die("SaveCheckpoint");
    smore_model_destroy(m);
    smore_graph_destroy(g);
    return 0;
}
