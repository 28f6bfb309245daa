// Host side of the bulk-exchange mode of the row-sharded store (device side: batch_kernels.cuh, "exchange" section).
//
// A super-batch moves three things between the shards, all as contiguous buffers:
//   REQ        requester s -> owner o : the local row ids s wants from o               (int32 lists)
//   ROWS_OUT   owner o -> requester s : those rows, in request order                   (rows into s's staging table)
//   ROWS_BACK  requester s -> owner o : the same rows after s's updates                (o adds back - sent to its table)
// The transport is NCCL send/recv groups (an all-to-all with per-peer counts) between the one-process-per-GPU ranks, or
// plain device-to-device copies when every shard lives in the calling process (tests: several shards on one device).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct smore_model_s;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);  // grows (contents are NOT preserved); SMORE_OK or an error code
    ~DevBuf() { cudaFree(p); }
};

// Everything one super-batch in flight needs; two sets alternate so that the exchange of super-batch i+1 overlaps the
// updates of super-batch i.
struct ExchSet {
    static constexpr int kW = 8;  // == kMaxWorld
    DevBuf hkey, hval, req, cnt /* [2][kW] int32: out, in */, off /* [kW] int32 */, wrk, req_in, sent, back;
    int cnt_out[kW] = {}, cnt_in[kW] = {};
    int64_t off_out[kW + 1] = {}, off_in[kW + 1] = {};
};

struct smore_exchange_s {
    static constexpr int kW = ExchSet::kW;
    int64_t superbatch = 0;       // samples per rank and super-batch (target)
    double hot_threshold = 0;     // a vertex expected to be a source >= this many times per super-batch (all ranks) is HOT
    int64_t n_hot = 0;
    DevBuf hot;                   // bitmap over vertex ids, null when n_hot == 0
    DevBuf errors;                // one int: samples the update kernel could not resolve (must stay 0)
    ExchSet set[2];
    int64_t req_stride = 0;
    // totals of the last train call
    uint64_t st_rows_moved = 0, st_superbatches = 0;
};

struct ExchTransport {
    enum { REQ = 0, ROWS_OUT = 1, ROWS_BACK = 2 };
    virtual ~ExchTransport() {}
    // device cnt_out of set b of every shard -> host cnt_out / cnt_in / off_out / off_in (synchronises `st` only)
    virtual int counts(smore_model_s** ms, int n, int b, cudaStream_t st) = 0;
    virtual int a2a(smore_model_s** ms, int n, int b, int what, size_t row_bytes, cudaStream_t st) = 0;
    // Collective: every rank passes the outcome of its set-up; returns SMORE_OK only when every rank passed SMORE_OK
    // (in-process shards: nothing to agree on).
    virtual int agree(int local_rc) { return local_rc; }
    // Called after a failed exchange: make queued transfers end so that the streams can be drained.
    virtual void abort() {}
};

ExchTransport* exch_local_transport();
ExchTransport* exch_nccl_transport(int rank, int world);  // nullptr (error set) unless smore_dist_nccl_init matched
