#include "train_line.inl"
template int train_line_t<float>(smore_model_s*, const smore_train_params*);
template int train_line_exchange_t<float>(smore_model_s**, int, const smore_train_params*, ExchTransport&);
template int train_mf_t<float>(smore_model_s*, const smore_train_params*);
template int train_line_block_t<float>(smore_model_s*, const smore_train_params*, int, void*, uint64_t);

// Debug hook: how many distinct SMs does a kernel launched on the carved-out update stream actually run on?
namespace {
__global__ void k_probe_smid(unsigned* bitmap, int spin) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) atomicOr(bitmap + (smid >> 5), 1u << (smid & 31));
    // keep the CTA alive for a while so that the whole grid has to spread over the partition
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
}
}  // namespace

extern "C" int smore_debug_sm_partition(int reserve, int* total_sms, int* partition_sms, int* sms_seen) {
    if (int rc = ensure_device()) return rc;
    setenv("SMORE_EXCH_RESERVE_SMS", std::to_string(reserve).c_str(), 1);
    ExchStreams* xsp = nullptr;
    if (int rc = exch_streams(&xsp)) return rc;
    ExchStreams& g_xs = *xsp;
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    unsigned* d = nullptr;
    CU(cudaMalloc((void**)&d, 64 * sizeof(unsigned)));
    CU(cudaMemset(d, 0, 64 * sizeof(unsigned)));
    k_probe_smid<<<sms * 8, 256, 0, g_xs.su>>>(d, 200000);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(g_xs.su));
    unsigned h[64];
    CU(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    int seen = 0;
    for (unsigned w : h) seen += __builtin_popcount(w);
    if (total_sms) *total_sms = sms;
    if (partition_sms) *partition_sms = g_xs.su_sms;
    if (sms_seen) *sms_seen = seen;
    return SMORE_OK;
}
