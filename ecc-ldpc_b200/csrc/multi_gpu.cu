// multi_gpu.cu -- single-process multi-GPU simulation for hosts that are ONE process (the Haskell tester):
// frames are independent, so GPU g of n decodes a contiguous range of the global frame index on its own replica of H,
// and the only exchange step is an NCCL sum all-reduce of the five u64 error counters over NVLink.
// (bench.py uses the other arrangement: one process per GPU + torch.distributed, same sharding.)
// NCCL is bound at run time with dlopen so that libldpc_b200.so has no link-time dependency on it.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

namespace ldpc {

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl()
{
    if (g_nccl.h) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("simulate_multi: cannot load libnccl.so.2 (%s)", dlerror()); return LDPC_B200_EUNSUPPORTED; }
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))dlsym(h, "ncclCommInitAll");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.CommInitAll || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        set_error("simulate_multi: libnccl lacks a required symbol");
        return LDPC_B200_EUNSUPPORTED;
    }
    g_nccl.h = h;
    return 0;
}

// The cached replicas are keyed by the CONTENT of the source code (dimensions + circulant table or CSR arrays), never
// by its address: a code destroyed and another one allocated at the same address must not hit a stale cache.
struct CodeKey {
    int M = 0, N = 0, sz = 0, brows = 0, bcols = 0;
    std::vector<int32_t> shifts, row_ptr, col_idx;
    bool valid = false;
    bool matches(const ldpc_b200_code *c) const
    {
        if (!valid || M != c->M || N != c->N || sz != c->sz || brows != c->brows || bcols != c->bcols) return false;
        return sz > 0 ? shifts == c->shifts : (row_ptr == c->row_ptr && col_idx == c->col_idx);
    }
    void take(const ldpc_b200_code *c)
    {
        M = c->M; N = c->N; sz = c->sz; brows = c->brows; bcols = c->bcols;
        if (sz > 0) { shifts = c->shifts; row_ptr.clear(); col_idx.clear(); }
        else { row_ptr = c->row_ptr; col_idx = c->col_idx; shifts.clear(); }
        valid = true;
    }
};

struct MultiCtx {
    CodeKey key;
    int n = 0, algo = 0, precision = 0, max_batch = 0;
    std::vector<int> devs;
    std::vector<ldpc_b200_code_t *> codes;
    std::vector<ldpc_b200_decoder_t *> decs;
    std::vector<uint64_t *> counters;
    std::vector<ncclComm_t> comms;
};

static MultiCtx g_ctx;
static std::mutex g_mu;

static void destroy_ctx()
{
    for (size_t g = 0; g < g_ctx.decs.size(); g++) {
        cudaSetDevice(g < g_ctx.devs.size() ? g_ctx.devs[g] : (int)g);
        if (g_ctx.decs[g]) ldpc_b200_decoder_destroy(g_ctx.decs[g]);
        if (g < g_ctx.codes.size() && g_ctx.codes[g]) ldpc_b200_code_destroy(g_ctx.codes[g]);
        if (g < g_ctx.counters.size() && g_ctx.counters[g]) cudaFree(g_ctx.counters[g]);
        if (g < g_ctx.comms.size() && g_ctx.comms[g]) g_nccl.CommDestroy(g_ctx.comms[g]);
    }
    g_ctx = MultiCtx();
}

static int build_ctx(const ldpc_b200_code *code, int n, int algo, int precision, int max_batch)
{
    if (g_ctx.key.matches(code) && g_ctx.n == n && g_ctx.algo == algo && g_ctx.precision == precision && g_ctx.max_batch == max_batch) return 0;
    destroy_ctx();
    g_ctx.key.take(code); g_ctx.n = n; g_ctx.algo = algo; g_ctx.precision = precision; g_ctx.max_batch = max_batch;
    g_ctx.codes.assign(n, nullptr); g_ctx.decs.assign(n, nullptr); g_ctx.counters.assign(n, nullptr); g_ctx.comms.assign(n, nullptr);
    g_ctx.devs.resize(n);
    for (int g = 0; g < n; g++) g_ctx.devs[g] = g;
    for (int g = 0; g < n; g++) {
        LDPC_CUDA(cudaSetDevice(g_ctx.devs[g]));
        int rc;
        if (code->sz > 0) rc = ldpc_b200_code_create_qc(code->sz, code->brows, code->bcols, code->shifts.data(), &g_ctx.codes[g]);
        else rc = ldpc_b200_code_create_csr(code->M, code->N, code->row_ptr.data(), code->col_idx.data(), &g_ctx.codes[g]);
        if (rc) return rc;
        rc = ldpc_b200_decoder_create(g_ctx.codes[g], algo, precision, max_batch, &g_ctx.decs[g]);
        if (rc) return rc;
        LDPC_CUDA(cudaMalloc(&g_ctx.counters[g], 8 * sizeof(uint64_t)));
    }
    if (n > 1) {
        int rc = load_nccl();
        if (rc) return rc;
        ncclResult_t r = g_nccl.CommInitAll(g_ctx.comms.data(), n, g_ctx.devs.data());
        if (r != ncclSuccess) { set_error("ncclCommInitAll failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return LDPC_B200_ECUDA; }
    }
    return 0;
}

}  // namespace ldpc

using namespace ldpc;

extern "C" int ldpc_b200_simulate_multi(const ldpc_b200_code_t *code, int n_gpus, int algo, int precision, int max_batch,
                                        int max_iters, int early_exit, double ebn0_db, int k_info, int n_tx, uint64_t seed,
                                        uint64_t first_frame, uint64_t n_frames, ldpc_b200_counters_t *out)
{
    if (!code || !out || n_gpus <= 0 || max_batch <= 0) { set_error("simulate_multi: bad arguments"); return LDPC_B200_EINVAL; }
    int avail = ldpc_b200_device_count();
    if (avail == 0) { set_error("no CUDA device available; libldpc_b200 has no CPU fallback"); return LDPC_B200_ENODEV; }
    if (n_gpus > avail) { set_error("simulate_multi: %d GPUs requested, %d visible", n_gpus, avail); return LDPC_B200_EINVAL; }
    std::lock_guard<std::mutex> lock(g_mu);
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = build_ctx(code, n_gpus, algo, precision, max_batch);
    if (rc) { destroy_ctx(); cudaSetDevice(prev); return rc; }
    // contiguous, balanced ranges of the global frame index
    std::vector<int> rcs(n_gpus, 0);
    std::vector<std::string> errs(n_gpus);
    uint64_t base = n_frames / n_gpus, rem = n_frames % n_gpus;
    auto worker = [&](int g) {
        uint64_t cnt = base + ((uint64_t)g < rem ? 1 : 0);
        uint64_t first = first_frame + (uint64_t)g * base + std::min<uint64_t>((uint64_t)g, rem);
        cudaSetDevice(g_ctx.devs[g]);
        cudaStream_t st = (cudaStream_t)ldpc_b200_decoder_stream(g_ctx.decs[g]);
        cudaMemsetAsync(g_ctx.counters[g], 0, 8 * sizeof(uint64_t), st);
        int r = cnt ? ldpc_b200_simulate_device(g_ctx.decs[g], max_iters, early_exit, ebn0_db, k_info, n_tx, seed, first, cnt, g_ctx.counters[g]) : 0;
        if (r) errs[g] = ldpc_b200_last_error();
        rcs[g] = r;
    };
    {
        std::vector<std::thread> th;
        for (int g = 1; g < n_gpus; g++) th.emplace_back(worker, g);
        worker(0);
        for (auto &t : th) t.join();
    }
    for (int g = 0; g < n_gpus; g++)
        if (rcs[g]) { set_error("simulate_multi: GPU %d: %s", g, errs[g].c_str()); cudaSetDevice(prev); return rcs[g]; }
    if (n_gpus > 1) {
        // the path's one exchange step: sum all-reduce of {frames, frame_errors, bit_errors, sum_iters, unconverged}
        g_nccl.GroupStart();
        for (int g = 0; g < n_gpus; g++) {
            cudaStream_t st = (cudaStream_t)ldpc_b200_decoder_stream(g_ctx.decs[g]);
            ncclResult_t r = g_nccl.AllReduce(g_ctx.counters[g], g_ctx.counters[g], 5, ncclUint64, ncclSum, g_ctx.comms[g], st);
            if (r != ncclSuccess) { g_nccl.GroupEnd(); set_error("ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); cudaSetDevice(prev); return LDPC_B200_ECUDA; }
        }
        g_nccl.GroupEnd();
    }
    double ms = 0.0;
    for (int g = 0; g < n_gpus; g++) {
        cudaSetDevice(g_ctx.devs[g]);
        rc = ldpc_b200_decoder_sync(g_ctx.decs[g]);
        double k = 0; int64_t nl = 0;
        if (!rc) rc = ldpc_b200_decoder_last_timing(g_ctx.decs[g], &k, &nl);
        if (rc) { cudaSetDevice(prev); return rc; }
        ms = std::max(ms, k);
    }
    uint64_t h[8];
    cudaSetDevice(g_ctx.devs[0]);
    LDPC_CUDA(cudaMemcpy(h, g_ctx.counters[0], sizeof h, cudaMemcpyDeviceToHost));
    cudaSetDevice(prev);
    out->frames += h[0]; out->frame_errors += h[1]; out->bit_errors += h[2]; out->sum_iters += h[3]; out->unconverged += h[4];
    out->kernel_ms += ms;
    return 0;
}

extern "C" void ldpc_b200_simulate_multi_release(void)
{
    std::lock_guard<std::mutex> lock(g_mu);
    destroy_ctx();
}
