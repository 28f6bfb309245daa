#include "host_graph.h"

#include <charconv>
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

int host_threads() {
    int t = (int)std::thread::hardware_concurrency();
    if (const char* e = std::getenv("SMORE_HOST_THREADS")) t = std::atoi(e);
    return std::max(1, std::min(t, 64));
}

double pow075(double x) {
    static const std::vector<double> lut = [] {
        std::vector<double> t(256);
        for (int k = 0; k < 256; ++k) t[(size_t)k] = std::pow((double)k, kPowerSample);
        return t;
    }();
    if (x >= 0.0 && x < 256.0) {
        const int k = (int)x;
        if ((double)k == x) return lut[(size_t)k];
    }
    return std::pow(x, kPowerSample);
}

void alias_method_cpp_into(const double* dist, int64_t n, double* prob, int64_t* alias, AliasScratch& sc) {
    // normalisation exactly as written in the reference: sum of pow, norm = n / sum, pow * norm (the reference evaluates
    // pow twice per entry, src/proNet.cpp:558,564; the value is the same, so it is computed once here)
    std::vector<double>& np = sc.np;
    np.resize((size_t)n);
    double sum = 0;
    for (int64_t i = 0; i < n; ++i) {
        np[(size_t)i] = pow075(dist[i]);
        sum += np[(size_t)i];
    }
    const double norm = (double)(size_t)n / sum;
    for (int64_t i = 0; i < n; ++i) {
        np[(size_t)i] = np[(size_t)i] * norm;
        prob[i] = 0.0;
        alias[i] = -1;
    }
    // two LIFO stacks filled in index order; the pop/pair/push order determines the table
    std::vector<int64_t>&small_block = sc.small_block, &large_block = sc.large_block;
    small_block.clear();
    large_block.clear();
    for (int64_t i = 0; i < n; ++i) {
        if (np[(size_t)i] < 1) small_block.push_back(i);
        else large_block.push_back(i);
    }
    while (!small_block.empty() && !large_block.empty()) {
        const int64_t s = small_block.back();
        small_block.pop_back();
        const int64_t l = large_block.back();
        large_block.pop_back();
        alias[s] = l;
        prob[s] = np[(size_t)s];
        np[(size_t)l] = np[(size_t)l] + np[(size_t)s] - 1;
        if (np[(size_t)l] < 1) small_block.push_back(l);
        else large_block.push_back(l);
    }
    for (int64_t l : large_block) prob[l] = 1.0;
    for (int64_t s : small_block) prob[s] = 1.0;
}

AliasHost alias_method_cpp(const double* dist, int64_t n) {
    AliasHost t;
    t.prob.resize((size_t)n);
    t.alias.resize((size_t)n);
    AliasScratch sc;
    alias_method_cpp_into(dist, n, t.prob.data(), t.alias.data(), sc);
    return t;
}

namespace {
// Fast paths of the two number formats for the magnitudes embeddings actually have. x * 10^k is formed with ONE rounding
// (10^|k| is exact in double for |k| <= 22), so its distance to the true value is <= half an ulp; whenever the scaled
// value is closer than `margin` (>> ulp) to a rounding boundary -- exact ties included -- the caller falls back to
// std::to_chars, which is correctly rounded. Both return the length written, 0 = not handled.
const double kP10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12};

inline char* put_digits(char* p, uint64_t v, int width) {  // zero-padded decimal
    for (int i = width - 1; i >= 0; --i) {
        p[i] = (char)('0' + v % 10);
        v /= 10;
    }
    return p + width;
}

int fast_g6(double x, char* out) {  // printf("%g")
    char* p = out;
    if (std::signbit(x)) *p++ = '-';
    const double ax = std::fabs(x);
    if (ax == 0.0) {
        *p++ = '0';
        return (int)(p - out);
    }
    if (!(ax >= 1e-7 && ax < 1e7)) return 0;  // also NaN
    int X = ax >= 1.0 ? 0 : -1;               // decimal exponent estimate, fixed up below
    if (ax >= 1.0) while (X < 6 && ax >= kP10[X + 1]) ++X;
    else while (X > -7 && ax * kP10[-X] < 1.0) --X;
    double y = X <= 5 ? ax * kP10[5 - X] : ax / kP10[X - 5];
    if (y < 1e5) { --X; y = X <= 5 ? ax * kP10[5 - X] : ax / kP10[X - 5]; }
    if (y >= 1e6) { ++X; y = X <= 5 ? ax * kP10[5 - X] : ax / kP10[X - 5]; }
    if (!(y >= 1e5 && y < 1e6) || X < -7 || X > 6) return 0;
    uint64_t n = (uint64_t)y;
    const double frac = y - (double)n;
    if (std::fabs(frac - 0.5) < 1e-6) return 0;  // too close to call: exact arithmetic decides
    if (frac > 0.5) ++n;
    if (n == 1000000) { n = 100000; ++X; }
    char d[6];
    put_digits(d, n, 6);
    int nd = 6;
    while (nd > 1 && d[nd - 1] == '0') --nd;  // %g strips trailing zeros
    if (X < -4 || X >= 6) {
        *p++ = d[0];
        if (nd > 1) {
            *p++ = '.';
            for (int i = 1; i < nd; ++i) *p++ = d[i];
        }
        *p++ = 'e';
        *p++ = X < 0 ? '-' : '+';
        p = put_digits(p, (uint64_t)(X < 0 ? -X : X), 2);
    } else if (X >= 0) {
        for (int i = 0; i <= X; ++i) *p++ = d[i];  // (nd may be shorter: the stripped zeros are still integer digits)
        if (nd > X + 1) {
            *p++ = '.';
            for (int i = X + 1; i < nd; ++i) *p++ = d[i];
        }
    } else {
        *p++ = '0';
        *p++ = '.';
        for (int i = 0; i < -X - 1; ++i) *p++ = '0';
        for (int i = 0; i < nd; ++i) *p++ = d[i];
    }
    return (int)(p - out);
}

int fast_f6(double x, char* out) {  // printf("%.6f")
    char* p = out;
    if (std::signbit(x)) *p++ = '-';
    const double ax = std::fabs(x);
    if (!(ax < 4096.0)) return 0;  // also NaN
    const double y = ax * 1e6;     // < 4.1e9: ulp <= 4.8e-7
    uint64_t n = (uint64_t)y;
    const double frac = y - (double)n;
    if (std::fabs(frac - 0.5) < 1e-5) return 0;
    if (frac > 0.5) ++n;
    const uint64_t ip = n / 1000000, fp = n % 1000000;
    char tmp[8];
    int ni = 0;
    uint64_t v = ip;
    do { tmp[ni++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (ni) *p++ = tmp[--ni];
    *p++ = '.';
    p = put_digits(p, fp, 6);
    return (int)(p - out);
}
}  // namespace

void format_rows(const double* rows, int64_t n, int dim, const std::string* names, int64_t first_id, int format,
                 std::string& out) {
    char num[400];  // "%.6f" of the largest double has 309 integer digits
    for (int64_t r = 0; r < n; ++r) {
        if (names) out += names[first_id + r];
        else {
            auto res = std::to_chars(num, num + sizeof(num), (long long)(first_id + r));
            out.append(num, (size_t)(res.ptr - num));
        }
        const double* row = rows + (size_t)r * (size_t)dim;
        for (int d = 0; d < dim; ++d) {
            num[0] = ' ';
            int len = format == 0 ? fast_g6(row[d], num + 1) : fast_f6(row[d], num + 1);
            if (len == 0) {
                // std::to_chars with a precision is specified as printf in the C locale: general/6 == "%g" (what
                // `ostream << double` prints at default precision), fixed/6 == "%.6f"
                auto res = format == 0 ? std::to_chars(num + 1, num + sizeof(num), row[d], std::chars_format::general, 6)
                                       : std::to_chars(num + 1, num + sizeof(num), row[d], std::chars_format::fixed, 6);
                len = (int)(res.ptr - (num + 1));
            }
            out.append(num, (size_t)len + 1);
        }
        out += '\n';
    }
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

namespace {
// Open-addressing string interner (ids in insertion order). Keys are views into the file buffer.
struct Interner {
    struct Slot { uint32_t hash; int32_t id; };
    std::vector<Slot> slots;
    std::vector<std::string_view> names;
    std::vector<uint32_t> hashes;  // hash of names[i]: re-used when the chunks are merged
    uint32_t mask = 0;
    explicit Interner(size_t expect = 1024) {
        size_t cap = 1024;
        while (cap < 2 * expect) cap <<= 1;
        slots.assign(cap, Slot{0, -1});
        mask = (uint32_t)(cap - 1);
    }
    static uint32_t hash_of(std::string_view s) {
        uint64_t h = 0xcbf29ce484222325ull;  // FNV-1a, folded
        for (unsigned char c : s) h = (h ^ c) * 0x100000001b3ull;
        return (uint32_t)(h ^ (h >> 32)) | 1u;
    }
    void grow() {
        std::vector<Slot> old;
        old.swap(slots);
        slots.assign(old.size() * 2, Slot{0, -1});
        mask = (uint32_t)(slots.size() - 1);
        for (const Slot& e : old)
            if (e.id >= 0) {
                uint32_t i = e.hash & mask;
                while (slots[i].id >= 0) i = (i + 1) & mask;
                slots[i] = e;
            }
    }
    int32_t intern(std::string_view s, uint32_t h) {
        uint32_t i = h & mask;
        while (slots[i].id >= 0) {
            if (slots[i].hash == h && names[(size_t)slots[i].id] == s) return slots[i].id;
            i = (i + 1) & mask;
        }
        const int32_t id = (int32_t)names.size();
        slots[i] = Slot{h, id};
        names.push_back(s);
        hashes.push_back(h);
        if (names.size() * 2 > slots.size()) grow();
        return id;
    }
};

// strtod semantics (what fscanf("%lf") / the reference accept) with std::from_chars doing the common cases
inline bool parse_weight(std::string_view t, double& w) {
    const char* b = t.data();
    const char* e = b + t.size();
    if (b < e && *b == '+') ++b;
    auto r = std::from_chars(b, e, w);
    if (r.ec == std::errc() && r.ptr == e) return true;
    std::string num(t);
    char* endp = nullptr;
    w = std::strtod(num.c_str(), &endp);
    return endp != num.c_str() && *endp == '\0';
}

struct Chunk {
    Interner names;
    std::vector<int32_t> src, dst;  // chunk-local ids
    std::vector<double> w;
    std::vector<int32_t> to_global;
};

void parse_chunk(const char* p, const char* end, Chunk& c) {
    auto is_space = [](char ch) { return ch == ' ' || ch == '\t' || ch == '\r' || ch == '\v' || ch == '\f'; };
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
        double w;
        if (nt == 3 && parse_weight(tok[2], w)) {
            const int32_t a = c.names.intern(tok[0], Interner::hash_of(tok[0]));  // first appearance: source before target
            const int32_t b = c.names.intern(tok[1], Interner::hash_of(tok[1]));
            c.src.push_back(a);
            c.dst.push_back(b);
            c.w.push_back(w);
        }
        p = eol + 1;
    }
}
}  // namespace

// The file is cut at line boundaries into one chunk per host thread. Each chunk is tokenised and interned locally (ids in
// first-appearance order INSIDE the chunk); the chunks' name lists are then merged in file order, which reproduces the
// global first-appearance order of the reference's sequential scan; finally the chunks rewrite their ids in parallel.
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

    const char* base = buf.data();
    const char* end = base + buf.size();
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), (int64_t)buf.size() / (1 << 20) + 1));
    std::vector<const char*> cut{base};
    for (int k = 1; k < nt; ++k) {
        const char* p = base + buf.size() / (size_t)nt * (size_t)k;
        if (p < cut.back()) p = cut.back();
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        cut.push_back(nl ? nl + 1 : end);
    }
    cut.push_back(end);
    std::vector<Chunk> chunks((size_t)nt);
    parallel_for(nt, 1, [&](int64_t b, int64_t e, int) {
        for (int64_t k = b; k < e; ++k) parse_chunk(cut[(size_t)k], cut[(size_t)k + 1], chunks[(size_t)k]);
    });
    // Merge, parallel over hash partitions: partition t interns every chunk's names whose hash falls to it, scanning the
    // chunks in file order, so each partition's list is in first-appearance order and remembers where (chunk, local id)
    // each name appeared first. A T-way merge of those lists by first appearance then assigns the global ids.
    struct First { int32_t chunk, local; };
    std::vector<Interner> part;
    part.reserve((size_t)nt);
    size_t uniq = 0;
    for (auto& c : chunks) uniq = std::max(uniq, c.names.names.size());
    for (int t = 0; t < nt; ++t) part.emplace_back(uniq * 2 / (size_t)nt + 1024);
    std::vector<std::vector<First>> first((size_t)nt);
    std::vector<int64_t> line0((size_t)nt + 1, 0);
    for (int k = 0; k < nt; ++k) {
        chunks[(size_t)k].to_global.resize(chunks[(size_t)k].names.names.size());
        line0[(size_t)k + 1] = line0[(size_t)k] + (int64_t)chunks[(size_t)k].src.size();
    }
    parallel_for(nt, 1, [&](int64_t tb, int64_t te, int) {
        for (int64_t t = tb; t < te; ++t) {
            Interner& in = part[(size_t)t];
            for (int k = 0; k < nt; ++k) {
                Chunk& c = chunks[(size_t)k];
                for (size_t i = 0; i < c.names.names.size(); ++i) {
                    const uint32_t h = c.names.hashes[i];
                    if ((int64_t)((h >> 8) % (uint32_t)nt) != t) continue;
                    const size_t before = in.names.size();
                    const int32_t pid = in.intern(c.names.names[i], h);
                    if (in.names.size() != before) first[(size_t)t].push_back(First{(int32_t)k, (int32_t)i});
                    c.to_global[i] = pid;  // partition-local for now (the partition is implied by the hash)
                }
            }
        }
    });
    std::vector<std::vector<int32_t>> gid((size_t)nt);  // partition-local id -> global id
    std::vector<std::string_view> global_names;
    {
        std::vector<size_t> head((size_t)nt, 0);
        size_t total = 0;
        for (int t = 0; t < nt; ++t) {
            gid[(size_t)t].resize(part[(size_t)t].names.size());
            total += part[(size_t)t].names.size();
        }
        global_names.reserve(total);
        for (size_t g = 0; g < total; ++g) {
            int best = -1;
            for (int t = 0; t < nt; ++t) {
                if (head[(size_t)t] >= first[(size_t)t].size()) continue;
                const First& a = first[(size_t)t][head[(size_t)t]];
                if (best < 0) { best = t; continue; }
                const First& b = first[(size_t)best][head[(size_t)best]];
                if (a.chunk < b.chunk || (a.chunk == b.chunk && a.local < b.local)) best = t;
            }
            gid[(size_t)best][head[(size_t)best]] = (int32_t)g;
            global_names.push_back(part[(size_t)best].names[head[(size_t)best]]);
            head[(size_t)best]++;
        }
    }
    parallel_for(nt, 1, [&](int64_t kb, int64_t ke, int) {
        for (int64_t k = kb; k < ke; ++k) {
            Chunk& c = chunks[(size_t)k];
            for (size_t i = 0; i < c.to_global.size(); ++i) {
                const uint32_t t = (c.names.hashes[i] >> 8) % (uint32_t)nt;
                c.to_global[i] = gid[t][(size_t)c.to_global[i]];
            }
        }
    });
    const int64_t V = (int64_t)global_names.size();
    const int64_t L = line0[(size_t)nt];
    std::vector<int32_t> src((size_t)L), dst((size_t)L);
    std::vector<double> wt((size_t)L);
    parallel_for(nt, 1, [&](int64_t b, int64_t e, int) {
        for (int64_t k = b; k < e; ++k) {
            const Chunk& c = chunks[(size_t)k];
            const int64_t o = line0[(size_t)k];
            for (size_t i = 0; i < c.src.size(); ++i) {
                src[(size_t)o + i] = c.to_global[(size_t)c.src[i]];
                dst[(size_t)o + i] = c.to_global[(size_t)c.dst[i]];
                wt[(size_t)o + i] = c.w[i];
            }
        }
    });
    out.n_lines = L;
    out.names.clear();
    out.names.resize((size_t)V);
    parallel_for(V, 1 << 14, [&](int64_t b, int64_t e, int) {
        for (int64_t v = b; v < e; ++v) out.names[(size_t)v].assign(global_names[(size_t)v]);
    });
    // CSR in insertion order: the forward entry, then (undirected) the reverse entry, line by line. Each host thread owns a
    // range of SOURCE vertices and scans all lines in file order, touching only its own vertices: adjacency order is the
    // file order, and the random accesses of a thread stay inside its slice.
    out.row_off.assign((size_t)V + 1, 0);
    const int ct = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), L / (1 << 18) + 1));
    auto vrange = [&](int t) { return std::pair<int64_t, int64_t>(V * t / ct, V * (t + 1) / ct); };
    parallel_for(ct, 1, [&](int64_t tb, int64_t te, int) {
        for (int64_t t = tb; t < te; ++t) {
            const auto [lo, hi] = vrange((int)t);
            for (int64_t i = 0; i < L; ++i) {
                const int64_t a = src[(size_t)i], b = dst[(size_t)i];
                if (a >= lo && a < hi) out.row_off[(size_t)a + 1]++;
                if (undirected && b >= lo && b < hi) out.row_off[(size_t)b + 1]++;
            }
        }
    });
    for (int64_t v = 0; v < V; ++v) out.row_off[(size_t)v + 1] += out.row_off[(size_t)v];
    const int64_t E = out.row_off[(size_t)V];
    out.col.assign((size_t)E, 0);
    out.w.assign((size_t)E, 0.0);
    std::vector<int64_t> cursor(out.row_off.begin(), out.row_off.end() - 1);
    parallel_for(ct, 1, [&](int64_t tb, int64_t te, int) {
        for (int64_t t = tb; t < te; ++t) {
            const auto [lo, hi] = vrange((int)t);
            for (int64_t i = 0; i < L; ++i) {
                const int64_t a = src[(size_t)i], b = dst[(size_t)i];
                if (a >= lo && a < hi) {
                    const int64_t e = cursor[(size_t)a]++;
                    out.col[(size_t)e] = (int32_t)b;
                    out.w[(size_t)e] = wt[(size_t)i];
                }
                if (undirected && b >= lo && b < hi) {
                    const int64_t e = cursor[(size_t)b]++;
                    out.col[(size_t)e] = (int32_t)a;
                    out.w[(size_t)e] = wt[(size_t)i];
                }
            }
        }
    });
    return true;
}

}  // namespace smore
