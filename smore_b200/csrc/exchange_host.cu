// Transports of the bulk-exchange mode (exchange_host.h) and the NCCL bootstrap entry points of the C ABI.
#include "host_common.h"

#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is dlopen()ed so that single-GPU hosts never need it

int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return SMORE_OK;
    cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4 + 256;  // head-room: the counts vary a little from one super-batch to the next
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return fail(SMORE_E_NOMEM, "cudaMalloc of %zu bytes for an exchange buffer failed: %s", want, cudaGetErrorString(e));
    cap = want;
    return SMORE_OK;
}

namespace {

constexpr int kW = smore_exchange_s::kW;

void prefix(ExchSet* x) {
    x->off_out[0] = x->off_in[0] = 0;
    for (int r = 0; r < kW; ++r) {
        x->off_out[r + 1] = x->off_out[r] + x->cnt_out[r];
        x->off_in[r + 1] = x->off_in[r] + x->cnt_in[r];
    }
}

// what -> (source buffer, element offset, element count, destination buffer on the peer, element offset there)
struct Leg {
    const char* src;
    char* dst;
    size_t bytes;
};

// The leg of `what` that goes from shard `a` to shard `b` (both in this process: local transport) -- or, with b == null,
// just the send half (src/bytes) and, with a == null, just the receive half (dst/bytes) as seen by the local shard.
Leg leg(int what, ExchSet* a, int64_t a_req_stride, int ra, ExchSet* b, int rb, size_t row_bytes) {
    Leg l{nullptr, nullptr, 0};
    switch (what) {
        case ExchTransport::REQ:  // requester a -> owner b
            if (a) { l.src = (const char*)a->req.p + ((size_t)rb * (size_t)a_req_stride) * 4; l.bytes = (size_t)a->cnt_out[rb] * 4; }
            if (b) { l.dst = (char*)b->req_in.p + (size_t)b->off_in[ra] * 4; l.bytes = (size_t)b->cnt_in[ra] * 4; }
            break;
        case ExchTransport::ROWS_OUT:  // owner a -> requester b
            if (a) { l.src = (const char*)a->sent.p + (size_t)a->off_in[rb] * row_bytes; l.bytes = (size_t)a->cnt_in[rb] * row_bytes; }
            if (b) { l.dst = (char*)b->wrk.p + (size_t)b->off_out[ra] * row_bytes; l.bytes = (size_t)b->cnt_out[ra] * row_bytes; }
            break;
        default:  // ROWS_BACK: requester a -> owner b
            if (a) { l.src = (const char*)a->wrk.p + (size_t)a->off_out[rb] * row_bytes; l.bytes = (size_t)a->cnt_out[rb] * row_bytes; }
            if (b) { l.dst = (char*)b->back.p + (size_t)b->off_in[ra] * row_bytes; l.bytes = (size_t)b->cnt_in[ra] * row_bytes; }
            break;
    }
    return l;
}

// ---- every shard in this process (tests: several shards on one device) -----------------------------------------------
struct LocalTransport : ExchTransport {
    int counts(smore_model_s** ms, int n, int b, cudaStream_t st) override {
        for (int r = 0; r < n; ++r) {
            ExchSet* x = &ms[r]->xch->set[b];
            memset(x->cnt_out, 0, sizeof(x->cnt_out));
            memset(x->cnt_in, 0, sizeof(x->cnt_in));
            CU(cudaMemcpyAsync(x->cnt_out, x->cnt.p, kW * sizeof(int), cudaMemcpyDeviceToHost, st));
        }
        CU(cudaStreamSynchronize(st));
        for (int r = 0; r < n; ++r)
            for (int s = 0; s < n; ++s) ms[r]->xch->set[b].cnt_in[s] = ms[s]->xch->set[b].cnt_out[r];
        for (int r = 0; r < n; ++r) prefix(&ms[r]->xch->set[b]);
        return SMORE_OK;
    }
    int a2a(smore_model_s** ms, int n, int b, int what, size_t row_bytes, cudaStream_t st) override {
        for (int a = 0; a < n; ++a)
            for (int c = 0; c < n; ++c) {
                if (a == c) continue;
                Leg l = leg(what, &ms[a]->xch->set[b], ms[a]->xch->req_stride, a, &ms[c]->xch->set[b], c, row_bytes);
                if (l.bytes) CU(cudaMemcpyAsync(l.dst, l.src, l.bytes, cudaMemcpyDeviceToDevice, st));
            }
        return SMORE_OK;
    }
};

// ---- one process per GPU: NCCL send/recv groups over NVLink ----------------------------------------------------------
struct NcclApi {
    void* so = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclCommAbort) CommAbort = nullptr;  // optional
    ncclComm_t comm = nullptr;
    int rank = -1, world = 0;
} g_nccl;

int nccl_load() {
    if (g_nccl.so) return SMORE_OK;
    // a host that already runs NCCL (torch.distributed) has libnccl.so.2 mapped: dlopen then returns that copy
    void* so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!so) return fail(SMORE_E_UNSUPPORTED, "libnccl.so.2 not found (%s): the bulk-exchange mode across processes needs NCCL", dlerror());
#define SYM(f)                                                                          \
    g_nccl.f = (decltype(g_nccl.f))dlsym(so, "nccl" #f);                                \
    if (!g_nccl.f) return fail(SMORE_E_UNSUPPORTED, "libnccl.so.2 lacks nccl" #f)
    SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(GroupStart); SYM(GroupEnd); SYM(Send); SYM(Recv);
    SYM(GetErrorString);
#undef SYM
    g_nccl.CommAbort = (decltype(g_nccl.CommAbort))dlsym(so, "ncclCommAbort");
    g_nccl.so = so;
    return SMORE_OK;
}

#define NC(call)                                                                                                        \
    do {                                                                                                                \
        ncclResult_t r_ = (call);                                                                                       \
        if (r_ != ncclSuccess) return fail(SMORE_E_CUDA, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

struct NcclTransport : ExchTransport {
    DevBuf status;  // [2][kW] int32: what I tell each peer, what each peer tells me
    int agree(int local_rc) override {
        // (a rank that could not even allocate these 64 bytes cannot take part: its peers then wait in the receive below
        // until the job's launcher notices the dead rank -- the one failure this protocol does not cover)
        if (int rc = status.ensure(2 * kW * sizeof(int))) return rc;
        if (!g_nccl.comm) return fail(SMORE_E_INVALID, "NCCL communicator was aborted by an earlier failure: call smore_dist_nccl_init again");
        int host[2 * kW];
        for (int r = 0; r < 2 * kW; ++r) host[r] = local_rc;
        CU(cudaMemcpy(status.p, host, sizeof(host), cudaMemcpyHostToDevice));
        int* d = (int*)status.p;
        NC(g_nccl.GroupStart());
        for (int r = 0; r < g_nccl.world; ++r) {
            if (r == g_nccl.rank) continue;
            NC(g_nccl.Send(d + r, 1, ncclInt32, r, g_nccl.comm, 0));
            NC(g_nccl.Recv(d + kW + r, 1, ncclInt32, r, g_nccl.comm, 0));
        }
        NC(g_nccl.GroupEnd());
        CU(cudaMemcpy(host, status.p, sizeof(host), cudaMemcpyDeviceToHost));
        if (local_rc) return local_rc;
        for (int r = 0; r < g_nccl.world; ++r)
            if (r != g_nccl.rank && host[kW + r] != SMORE_OK)
                return fail(host[kW + r], "exchange mode: rank %d failed to set up (%s); no rank trains", r, "see that rank's smore_last_error");
        return SMORE_OK;
    }
    void abort() override {
        if (!g_nccl.comm) return;
        if (g_nccl.CommAbort) g_nccl.CommAbort(g_nccl.comm);
        else g_nccl.CommDestroy(g_nccl.comm);
        g_nccl.comm = nullptr;
        g_nccl.rank = -1;
        g_nccl.world = 0;
    }
    int counts(smore_model_s** ms, int n, int b, cudaStream_t st) override {
        if (n != 1) return fail(SMORE_E_INVALID, "the NCCL transport drives exactly one shard per process");
        ExchSet* x = &ms[0]->xch->set[b];
        int* d_out = (int*)x->cnt.p;
        int* d_in = d_out + kW;
        const int me = g_nccl.rank;
        NC(g_nccl.GroupStart());
        for (int r = 0; r < g_nccl.world; ++r) {
            if (r == me) continue;
            NC(g_nccl.Send(d_out + r, 1, ncclInt32, r, g_nccl.comm, st));
            NC(g_nccl.Recv(d_in + r, 1, ncclInt32, r, g_nccl.comm, st));
        }
        NC(g_nccl.GroupEnd());
        int both[2 * kW];
        CU(cudaMemcpyAsync(both, d_out, sizeof(both), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        memset(x->cnt_out, 0, sizeof(x->cnt_out));
        memset(x->cnt_in, 0, sizeof(x->cnt_in));
        for (int r = 0; r < g_nccl.world; ++r)
            if (r != me) { x->cnt_out[r] = both[r]; x->cnt_in[r] = both[kW + r]; }
        prefix(x);
        return SMORE_OK;
    }
    int a2a(smore_model_s** ms, int n, int b, int what, size_t row_bytes, cudaStream_t st) override {
        if (n != 1) return fail(SMORE_E_INVALID, "the NCCL transport drives exactly one shard per process");
        ExchSet* x = &ms[0]->xch->set[b];
        const int64_t stride = ms[0]->xch->req_stride;
        const int me = g_nccl.rank;
        NC(g_nccl.GroupStart());
        for (int r = 0; r < g_nccl.world; ++r) {
            if (r == me) continue;
            Leg s = leg(what, x, stride, me, nullptr, r, row_bytes);  // what I send to r
            Leg d = leg(what, nullptr, 0, r, x, me, row_bytes);       // what I receive from r
            if (s.bytes) NC(g_nccl.Send(s.src, s.bytes, ncclInt8, r, g_nccl.comm, st));
            if (d.bytes) NC(g_nccl.Recv(d.dst, d.bytes, ncclInt8, r, g_nccl.comm, st));
        }
        NC(g_nccl.GroupEnd());
        return SMORE_OK;
    }
};

LocalTransport g_local;
NcclTransport g_nccl_tr;

}  // namespace

ExchTransport* exch_local_transport() { return &g_local; }

ExchTransport* exch_nccl_transport(int rank, int world) {
    if (!g_nccl.comm) {
        fail(SMORE_E_INVALID, "bulk-exchange mode across processes: call smore_dist_nccl_init first");
        return nullptr;
    }
    if (g_nccl.rank != rank || g_nccl.world != world) {
        fail(SMORE_E_INVALID, "NCCL communicator is rank %d of %d but the graph shard is rank %d of %d", g_nccl.rank, g_nccl.world, rank, world);
        return nullptr;
    }
    return &g_nccl_tr;
}

extern "C" {

int smore_dist_nccl_unique_id(void* id128) {
    if (!id128) return fail(SMORE_E_INVALID, "null buffer");
    if (int rc = nccl_load()) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return SMORE_OK;
}

int smore_dist_nccl_init(const void* id128, int rank, int world) {
    if (!id128 || world < 2 || world > kW || rank < 0 || rank >= world) return fail(SMORE_E_INVALID, "bad NCCL bootstrap arguments");
    if (int rc = nccl_load()) return rc;
    if (int rc = ensure_device()) return rc;
    if (g_nccl.comm) return fail(SMORE_E_INVALID, "NCCL communicator already initialised");
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NC(g_nccl.CommInitRank(&g_nccl.comm, world, id, rank));
    g_nccl.rank = rank;
    g_nccl.world = world;
    return SMORE_OK;
}

int smore_dist_nccl_shutdown(void) {
    if (g_nccl.comm) {
        cudaDeviceSynchronize();
        g_nccl.CommDestroy(g_nccl.comm);
        g_nccl.comm = nullptr;
        g_nccl.rank = -1;
        g_nccl.world = 0;
    }
    return SMORE_OK;
}

}  // extern "C"
