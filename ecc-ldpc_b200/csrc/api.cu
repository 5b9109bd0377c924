// api.cu -- the C ABI of libldpc_b200.so (include/ldpc_b200.h). Host-side glue only; kernels live in
// stream_engine.cu (any H, HBM-streaming), resident_engine.cu (quasi-cyclic, on-chip) and channel.cu.
#include "common.cuh"
#include "philox.cuh"
namespace ldpc {
#include "resident_common.cuh"
}
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <set>
#include <thread>
#include <new>

namespace ldpc {

static thread_local char g_err[512] = "";
static int g_device = -1;
static bool g_init = false;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    if (e == cudaErrorMemoryAllocation) return LDPC_B200_ENOMEM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return LDPC_B200_ENODEV;
    return LDPC_B200_ECUDA;
}

// Live handles: ldpc_b200_shutdown (the Code's `finalize`, Utils.hs:91-103) releases whatever the host never destroyed,
// and the destroy functions ignore a handle that is no longer registered (already released, e.g. by shutdown).
static std::mutex g_reg_mu;
static std::set<ldpc_b200_decoder *> g_live_decs;
static std::set<ldpc_b200_code *> g_live_codes;
template <class H> static void reg_add(std::set<H *> &s, H *h) { std::lock_guard<std::mutex> lk(g_reg_mu); s.insert(h); }
template <class H> static bool reg_remove(std::set<H *> &s, H *h) { std::lock_guard<std::mutex> lk(g_reg_mu); return s.erase(h) != 0; }

// One device copy of H per (device, content): mkLDPC calls `decoder0 h` maxThreadCount times (Utils.hs:53), i.e. a binding
// creates the same code once per tester thread.  Sharing the object saves the copies and, more importantly, makes all those
// decoder handles meet in ONE mailbox (mailbox.cu), which is what lets their one-frame calls coalesce.
static std::mutex g_create_mu;
static ldpc_b200_code *find_shared_code(int device, int M, int N, int sz, int brows, int bcols, const int32_t *shifts,
                                        const int32_t *row_ptr, const int32_t *col_idx)
{
    std::lock_guard<std::mutex> lk(g_reg_mu);
    for (ldpc_b200_code *c : g_live_codes) {
        if (c->device != device || c->M != M || c->N != N || c->sz != sz || c->brows != brows || c->bcols != bcols) continue;
        const bool same = sz > 0 ? std::equal(c->shifts.begin(), c->shifts.end(), shifts)
                                 : (std::equal(c->row_ptr.begin(), c->row_ptr.end(), row_ptr) &&
                                    std::equal(c->col_idx.begin(), c->col_idx.begin() + c->row_ptr[M], col_idx));
        if (same) { c->refs++; return c; }
    }
    return nullptr;
}

static int require_init()
{
    if (!g_init) {
        int rc = ldpc_b200_init(-1);
        if (rc) return rc;
    }
    return 0;
}

static int upload_graph(ldpc_b200_code *c)
{
    DeviceGraph &g = c->g;
    g.M = c->M; g.N = c->N; g.E = c->E; g.max_row_deg = c->max_row_deg; g.max_col_deg = c->max_col_deg;
    size_t e = c->E > 0 ? c->E : 1;
    LDPC_CUDA(cudaMalloc(&g.row_ptr, sizeof(int32_t) * (c->M + 1)));
    LDPC_CUDA(cudaMalloc(&g.col_idx, sizeof(int32_t) * e));
    LDPC_CUDA(cudaMalloc(&g.col_ptr, sizeof(int32_t) * (c->N + 1)));
    LDPC_CUDA(cudaMalloc(&g.col_edge, sizeof(int32_t) * e));
    LDPC_CUDA(cudaMemcpy(g.row_ptr, c->row_ptr.data(), sizeof(int32_t) * (c->M + 1), cudaMemcpyHostToDevice));
    LDPC_CUDA(cudaMemcpy(g.col_ptr, c->col_ptr.data(), sizeof(int32_t) * (c->N + 1), cudaMemcpyHostToDevice));
    if (c->E > 0) {
        LDPC_CUDA(cudaMemcpy(g.col_idx, c->col_idx.data(), sizeof(int32_t) * c->E, cudaMemcpyHostToDevice));
        LDPC_CUDA(cudaMemcpy(g.col_edge, c->col_edge.data(), sizeof(int32_t) * c->E, cudaMemcpyHostToDevice));
    }
    // fixed-stride records (see DeviceGraph)
    const int classes[4] = {4, 8, 18, 32};
    g.row_ell_d = 0;
    for (int k = 0; k < 4 && !g.row_ell_d; k++) if (c->max_row_deg <= classes[k]) g.row_ell_d = classes[k];
    if (g.row_ell_d && c->M > 0) {
        const int S = g.row_ell_d + 2;
        std::vector<int32_t> t((size_t)c->M * S, 0);
        for (int m = 0; m < c->M; m++) {
            int b = c->row_ptr[m], d = c->row_ptr[m + 1] - b;
            t[(size_t)m * S] = b; t[(size_t)m * S + 1] = d;
            for (int k = 0; k < g.row_ell_d; k++) t[(size_t)m * S + 2 + k] = d > 0 ? c->col_idx[b + (k < d ? k : d - 1)] : 0;
        }
        LDPC_CUDA(cudaMalloc(&g.row_ell, sizeof(int32_t) * t.size()));
        LDPC_CUDA(cudaMemcpy(g.row_ell, t.data(), sizeof(int32_t) * t.size(), cudaMemcpyHostToDevice));
    }
    {
        const int D = DeviceGraph::COL_ELL_D, S = D + 1;
        std::vector<int32_t> t((size_t)c->N * S, 0);
        for (int j = 0; j < c->N; j++) {
            int q0 = c->col_ptr[j], d = c->col_ptr[j + 1] - q0;
            t[(size_t)j * S] = d;
            for (int k = 0; k < D; k++) t[(size_t)j * S + 1 + k] = d > 0 ? c->col_edge[q0 + d - 1 - (k < d ? k : d - 1)] : 0;
        }
        LDPC_CUDA(cudaMalloc(&g.col_ell, sizeof(int32_t) * t.size()));
        LDPC_CUDA(cudaMemcpy(g.col_ell, t.data(), sizeof(int32_t) * t.size(), cudaMemcpyHostToDevice));
    }
    if (g.row_ell_d && c->E > 0) {
        int rc = onchip_build_tables(c);          // shared-memory placement + records of the ONCHIP engine (may decline)
        if (rc) return rc;
    }
    return 0;
}

// Fills the CSC side and the degree maxima from row_ptr / col_idx.
static int finish_code(ldpc_b200_code *c)
{
    int M = c->M, N = c->N;
    c->E = c->row_ptr[M];
    c->col_ptr.assign(N + 1, 0);
    c->col_edge.assign(c->E > 0 ? c->E : 1, 0);
    c->max_row_deg = 0;
    for (int m = 0; m < M; m++) {
        int d = c->row_ptr[m + 1] - c->row_ptr[m];
        c->max_row_deg = std::max(c->max_row_deg, d);
        for (int e = c->row_ptr[m]; e < c->row_ptr[m + 1]; e++) c->col_ptr[c->col_idx[e] + 1]++;
    }
    c->max_col_deg = 0;
    for (int j = 0; j < N; j++) {
        c->max_col_deg = std::max(c->max_col_deg, c->col_ptr[j + 1]);
        c->col_ptr[j + 1] += c->col_ptr[j];
    }
    std::vector<int32_t> fill(N, 0);
    for (int m = 0; m < M; m++)
        for (int e = c->row_ptr[m]; e < c->row_ptr[m + 1]; e++) {
            int j = c->col_idx[e];
            c->col_edge[c->col_ptr[j] + fill[j]++] = e;
        }
    return 0;
}

static int pick_engine(const ldpc_b200_decoder *d)
{
    if (d->engine == LDPC_B200_ENGINE_STREAM) return LDPC_B200_ENGINE_STREAM;
    if (resident_supports(d)) return LDPC_B200_ENGINE_RESIDENT;
    return LDPC_B200_ENGINE_STREAM;
}

static int ensure_staging(ldpc_b200_decoder *d)
{
    if (d->d_bits) return 0;
    int N = d->code->N;
    size_t B = (size_t)d->max_batch;
    LDPC_CUDA(cudaMalloc(&d->d_in, B * N * sizeof(double)));
    LDPC_CUDA(cudaMalloc(&d->d_bits, B * words_per_frame(N) * sizeof(uint32_t)));
    LDPC_CUDA(cudaMalloc(&d->d_iters, B * sizeof(int32_t)));
    LDPC_CUDA(cudaMalloc(&d->d_conv, B));
    LDPC_CUDA(cudaMalloc(&d->d_counters, 8 * sizeof(uint64_t)));
    return 0;
}

static int decode_dispatch(ldpc_b200_decoder *d, int max_iters, int early_exit, int n_frames, const void *d_llr,
                           bool llr_is_double, uint32_t *d_bits, int32_t *d_iters, uint8_t *d_conv,
                           int src_stride = 0, int src_take = 0)
{
    if (pick_engine(d) == LDPC_B200_ENGINE_RESIDENT)
        return resident_decode(d, max_iters, early_exit, n_frames, d_llr, llr_is_double, d_bits, d_iters, d_conv, src_stride, src_take);
    return stream_decode(d, max_iters, early_exit, n_frames, d_llr, llr_is_double, d_bits, d_iters, d_conv, nullptr, nullptr);
}

static void begin_timing(ldpc_b200_decoder *d)
{
    d->launches = 0;
    cudaEventRecord(d->ev0, d->stream);
}

static void end_timing(ldpc_b200_decoder *d)
{
    cudaEventRecord(d->ev1, d->stream);
    d->timing_pending = true;
    d->last_launches = d->launches;
}

static int finish_timing(ldpc_b200_decoder *d)
{
    if (d->timing_pending) {
        LDPC_CUDA(cudaEventSynchronize(d->ev1));
        float ms = 0.f;
        LDPC_CUDA(cudaEventElapsedTime(&ms, d->ev0, d->ev1));
        d->last_ms = ms;
        d->timing_pending = false;
    }
    return 0;
}

static void channel_params(double ebn0_db, int k_info, int n_tx, float *sigma, float *scale)
{
    // sigma^2 = 1 / (2 R Eb/N0), R = k_info / n_tx (pinned statistically by NOTES.txt:3,8,13)
    double R = (double)k_info / (double)n_tx;
    double s2 = 1.0 / (2.0 * R * pow(10.0, ebn0_db / 10.0));
    *sigma = (float)sqrt(s2);
    *scale = (float)(2.0 / s2);
}

// decode_trace on the RESIDENT / ONCHIP engines: one frame, messages of every executed iteration in reference edge order
template <typename T>
static int trace_onchip_engines_t(ldpc_b200_decoder *d, int max_iters, const double *llr, double *c2v, double *lam)
{
    const ldpc_b200_code *c = d->code;
    const int N = c->N, E = c->E, kind = resident_kind_of(d);
    const int NE = kind == 3 ? c->g.oc_se : E;                    // ONCHIP dumps its message slots
    const size_t rows = (size_t)std::max(max_iters, 1);
    std::vector<T> h((size_t)N);
    for (int j = 0; j < N; j++) h[j] = (T)llr[j];                 // the rounding the batched path applies on the device
    T *d_c2v = nullptr, *d_lam = nullptr;
    cudaError_t e = cudaMalloc(&d_c2v, rows * NE * sizeof(T));
    if (e == cudaSuccess) e = cudaMalloc(&d_lam, rows * N * sizeof(T));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d->d_in, h.data(), (size_t)N * sizeof(T), cudaMemcpyHostToDevice, d->stream);
    int rc = e == cudaSuccess ? 0 : cuda_fail(e, "decode_trace: buffers", __FILE__, __LINE__);
    if (!rc) {
        begin_timing(d);
        if (kind == 3) {
            ResidentParams p = {};
            p.llr = d->d_in; p.bits = d->d_bits; p.iters = d->d_iters; p.conv = d->d_conv;
            p.n_frames = 1; p.max_iters = max_iters; p.early_exit = 1; p.n_tx = N;
            p.trace_c2v = d_c2v; p.trace_lam = d_lam; p.trace_ne = NE;
            rc = onchip_launch(d, p, d->stream, true);
        } else {
            rc = resident_trace(d, max_iters, d->d_in, d->d_bits, d->d_iters, d->d_conv, d_c2v, d_lam);
        }
        end_timing(d);
    }
    int32_t it = 0;
    if (!rc) {
        e = cudaMemcpyAsync(&it, d->d_iters, sizeof it, cudaMemcpyDeviceToHost, d->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "decode_trace: kernel", __FILE__, __LINE__);
    }
    if (!rc && it > 0) {
        std::vector<T> hc((size_t)it * NE), hl((size_t)it * N);
        e = cudaMemcpy(hc.data(), d_c2v, hc.size() * sizeof(T), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(hl.data(), d_lam, hl.size() * sizeof(T), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "decode_trace: copy back", __FILE__, __LINE__);
        for (int t = 0; t < it && !rc; t++) {
            if (c2v) for (int q = 0; q < E; q++) c2v[(size_t)t * E + q] = (double)hc[(size_t)t * NE + (kind == 3 ? c->oc_slot_of_edge[q] : q)];
            if (lam) for (int j = 0; j < N; j++) lam[(size_t)t * N + j] = (double)hl[(size_t)t * N + (kind == 3 ? c->oc_pos_of_var[j] : j)];
        }
    }
    cudaFree(d_c2v); cudaFree(d_lam);
    return rc;
}

static int trace_onchip_engines(ldpc_b200_decoder *d, int max_iters, const double *llr, double *c2v, double *lam)
{
    return d->precision == LDPC_B200_PREC_FP64 ? trace_onchip_engines_t<double>(d, max_iters, llr, c2v, lam)
                                               : trace_onchip_engines_t<float>(d, max_iters, llr, c2v, lam);
}

}  // namespace ldpc

using namespace ldpc;

extern "C" {

int ldpc_b200_abi_version(void) { return LDPC_B200_ABI_VERSION; }
const char *ldpc_b200_last_error(void) { return g_err; }

int ldpc_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ldpc_b200_init(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libldpc_b200 has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        return LDPC_B200_ENODEV;
    }
    if (device >= n) { set_error("device %d out of range (%d devices)", device, n); return LDPC_B200_EINVAL; }
    if (device >= 0) LDPC_CUDA(cudaSetDevice(device));
    LDPC_CUDA(cudaGetDevice(&g_device));
    LDPC_CUDA(cudaFree(0));
    g_init = true;
    return 0;
}

void ldpc_b200_shutdown(void)
{
    // everything the library still holds on the device goes: the replicas of ldpc_b200_simulate_multi, then every live
    // decoder (streams, staging, pinned buffers), then every live code (device copies of H)
    ldpc_b200_simulate_multi_release();
    for (;;) {
        ldpc_b200_decoder *d = nullptr;
        { std::lock_guard<std::mutex> lk(g_reg_mu); if (!g_live_decs.empty()) d = *g_live_decs.begin(); }
        if (!d) break;
        ldpc_b200_decoder_destroy(d);
    }
    for (;;) {
        ldpc_b200_code *c = nullptr;
        { std::lock_guard<std::mutex> lk(g_reg_mu); if (!g_live_codes.empty()) { c = *g_live_codes.begin(); c->refs = 1; c->dec_refs = 0; } }
        if (!c) break;
        ldpc_b200_code_destroy(c);
    }
    g_init = false;
}

int ldpc_b200_live_handles(int *codes, int *decoders)
{
    std::lock_guard<std::mutex> lk(g_reg_mu);
    if (codes) { *codes = 0; for (const ldpc_b200_code *c : g_live_codes) *codes += c->refs; }      // handles, not objects
    if (decoders) *decoders = (int)g_live_decs.size();
    return 0;
}

int ldpc_b200_code_create_csr(int M, int N, const int32_t *row_ptr, const int32_t *col_idx, ldpc_b200_code_t **out)
{
    if (!out || M < 0 || N <= 0 || !row_ptr || (!col_idx && row_ptr[M] > 0)) { set_error("code_create_csr: bad arguments"); return LDPC_B200_EINVAL; }
    if (row_ptr[0] != 0) { set_error("code_create_csr: row_ptr[0] must be 0"); return LDPC_B200_EINVAL; }
    for (int m = 0; m < M; m++) {
        if (row_ptr[m + 1] < row_ptr[m]) { set_error("code_create_csr: row_ptr not monotone at row %d", m); return LDPC_B200_EINVAL; }
        for (int e = row_ptr[m]; e < row_ptr[m + 1]; e++) {
            if (col_idx[e] < 0 || col_idx[e] >= N) { set_error("code_create_csr: column %d out of range in row %d", col_idx[e], m); return LDPC_B200_EINVAL; }
            if (e > row_ptr[m] && col_idx[e] <= col_idx[e - 1]) { set_error("code_create_csr: columns of row %d not strictly ascending", m); return LDPC_B200_EINVAL; }
        }
    }
    int rc = require_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> create_lock(g_create_mu);
    {
        int dev = -1;
        LDPC_CUDA(cudaGetDevice(&dev));
        if (ldpc_b200_code *shared = find_shared_code(dev, M, N, 0, 0, 0, nullptr, row_ptr, col_idx)) { *out = shared; return 0; }
    }
    ldpc_b200_code *c = new (std::nothrow) ldpc_b200_code();
    if (!c) return LDPC_B200_ENOMEM;
    c->M = M; c->N = N;
    c->row_ptr.assign(row_ptr, row_ptr + M + 1);
    c->col_idx.assign(col_idx, col_idx + row_ptr[M]);
    finish_code(c);
    {   // the code lives on the CURRENT device (multi-GPU hosts switch devices)
        cudaError_t ce = cudaGetDevice(&c->device);
        if (ce != cudaSuccess) { delete c; return cuda_fail(ce, "cudaGetDevice", __FILE__, __LINE__); }
    }
    reg_add(g_live_codes, c);
    rc = upload_graph(c);
    if (rc) { ldpc_b200_code_destroy(c); return rc; }
    if (onchip_fits(c, LDPC_B200_PREC_FP32)) c->resident_kind = 3;      // run-time table on-chip kernel
    *out = c;
    return 0;
}

int ldpc_b200_code_create_qc(int sz, int brows, int bcols, const int32_t *shifts, ldpc_b200_code_t **out)
{
    if (!out || sz <= 0 || brows <= 0 || bcols <= 0 || !shifts) { set_error("code_create_qc: bad arguments"); return LDPC_B200_EINVAL; }
    for (int i = 0; i < brows * bcols; i++)
        if (shifts[i] < -1 || shifts[i] >= sz) { set_error("code_create_qc: shift %d out of range [-1, %d)", shifts[i], sz); return LDPC_B200_EINVAL; }
    int rc = require_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> create_lock(g_create_mu);
    {
        int dev = -1;
        LDPC_CUDA(cudaGetDevice(&dev));
        if (ldpc_b200_code *shared = find_shared_code(dev, brows * sz, bcols * sz, sz, brows, bcols, shifts, nullptr, nullptr)) { *out = shared; return 0; }
    }
    ldpc_b200_code *c = new (std::nothrow) ldpc_b200_code();
    if (!c) return LDPC_B200_ENOMEM;
    c->M = brows * sz; c->N = bcols * sz;
    c->sz = sz; c->brows = brows; c->bcols = bcols;
    c->shifts.assign(shifts, shifts + brows * bcols);
    // QuasiCyclic.hs:19-25: row r*sz+i has a one in column c*sz + (i+k) mod sz for every block with shift k.
    c->row_ptr.assign(c->M + 1, 0);
    for (int r = 0; r < brows; r++)
        for (int i = 0; i < sz; i++) {
            int m = r * sz + i;
            for (int b = 0; b < bcols; b++) {
                int k = shifts[r * bcols + b];
                if (k >= 0) c->col_idx.push_back(b * sz + (i + k) % sz);
            }
            c->row_ptr[m + 1] = (int32_t)c->col_idx.size();
        }
    finish_code(c);
    {
        cudaError_t ce = cudaGetDevice(&c->device);
        if (ce != cudaSuccess) { delete c; return cuda_fail(ce, "cudaGetDevice", __FILE__, __LINE__); }
    }
    c->resident_kind = resident_kind_for(c);
    reg_add(g_live_codes, c);
    rc = upload_graph(c);
    if (rc) { ldpc_b200_code_destroy(c); return rc; }
    if (!c->resident_kind && onchip_fits(c, LDPC_B200_PREC_FP32)) c->resident_kind = 3;
    *out = c;
    return 0;
}

void ldpc_b200_code_destroy(ldpc_b200_code_t *c)
{
    if (!c) return;
    {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        auto it = g_live_codes.find(c);
        if (it == g_live_codes.end()) return;              // already released (e.g. by ldpc_b200_shutdown)
        if (c->refs > 0) c->refs--;
        if (c->refs > 0 || c->dec_refs > 0) return;         // other handles share it, or decoders still decode with it
        g_live_codes.erase(it);
    }
    cudaFree(c->g.row_ptr); cudaFree(c->g.col_idx); cudaFree(c->g.col_ptr); cudaFree(c->g.col_edge);
    cudaFree(c->g.row_ell); cudaFree(c->g.col_ell); cudaFree(c->g.oc_tab); cudaFree(c->g.oc_posvar); cudaFree(c->g.col_slot);
    delete c;
}

int ldpc_b200_code_info(const ldpc_b200_code_t *c, int *M, int *N, int *E, int *qc_sz)
{
    if (!c) { set_error("code_info: null code"); return LDPC_B200_EINVAL; }
    if (M) *M = c->M;
    if (N) *N = c->N;
    if (E) *E = c->E;
    if (qc_sz) *qc_sz = c->sz;
    return 0;
}

int ldpc_b200_decoder_create(const ldpc_b200_code_t *code, int algo, int precision, int max_batch, ldpc_b200_decoder_t **out)
{
    if (!code || !out || max_batch <= 0) { set_error("decoder_create: bad arguments"); return LDPC_B200_EINVAL; }
    if (algo < LDPC_B200_ALGO_SPA || algo > LDPC_B200_ALGO_MINSUM_ARRAYLET) { set_error("decoder_create: unknown algo %d", algo); return LDPC_B200_EINVAL; }
    if (algo == LDPC_B200_ALGO_SPA_CM && precision != LDPC_B200_PREC_FP64) {
        set_error("decoder_create: the cached-product sum-product (CachedMult.hs) is a float64 verification variant; use LDPC_B200_ALGO_SPA for float32");
        return LDPC_B200_EUNSUPPORTED;
    }
    if (precision != LDPC_B200_PREC_FP64 && precision != LDPC_B200_PREC_FP32) { set_error("decoder_create: unknown precision %d", precision); return LDPC_B200_EINVAL; }
    int rc = require_init();
    if (rc) return rc;
    ldpc_b200_decoder *d = new (std::nothrow) ldpc_b200_decoder();
    if (!d) return LDPC_B200_ENOMEM;
    d->code = code; d->algo = algo; d->precision = precision; d->max_batch = max_batch;
    d->device = code->device;
    { std::lock_guard<std::mutex> lk(g_reg_mu); const_cast<ldpc_b200_code *>(code)->dec_refs++; }
    reg_add(g_live_decs, d);
    cudaError_t ce = cudaSetDevice(code->device);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreate(&d->ev0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&d->ev1);
    if (ce != cudaSuccess) {                      // nothing half-built escapes: destroy tolerates the null members
        ldpc_b200_decoder_destroy(d);
        return cuda_fail(ce, "decoder_create: stream / event creation", __FILE__, __LINE__);
    }
    *out = d;
    return 0;
}

int ldpc_b200_decoder_set_engine(ldpc_b200_decoder_t *d, int engine)
{
    if (!d || engine < 0 || engine > 3) { set_error("set_engine: bad arguments"); return LDPC_B200_EINVAL; }
    if (engine == LDPC_B200_ENGINE_RESIDENT && d->code->resident_kind != 1 && d->code->resident_kind != 2) {
        set_error("RESIDENT engine has no compiled specialisation for this code (ONCHIP is the run-time table form)");
        return LDPC_B200_EUNSUPPORTED;
    }
    int prev = d->engine;
    d->engine = engine;
    if ((engine == LDPC_B200_ENGINE_RESIDENT || engine == LDPC_B200_ENGINE_ONCHIP) && !resident_supports(d)) {
        d->engine = prev;
        set_error("no on-chip kernel for this code / algorithm / precision");
        return LDPC_B200_EUNSUPPORTED;
    }
    return 0;
}

int ldpc_b200_decoder_engine(const ldpc_b200_decoder_t *d)
{
    if (!d) return LDPC_B200_EINVAL;
    int e = pick_engine(d);
    return (e == LDPC_B200_ENGINE_RESIDENT && resident_kind_of(d) == 3) ? LDPC_B200_ENGINE_ONCHIP : e;
}

void ldpc_b200_decoder_destroy(ldpc_b200_decoder_t *d)
{
    if (!d || !reg_remove(g_live_decs, d)) return;
    cudaSetDevice(d->device);
    mailbox_release(d);
    if (d->stream) cudaStreamSynchronize(d->stream);
    stream_free(d);
    cudaFree(d->d_in); cudaFree(d->d_bits); cudaFree(d->d_iters); cudaFree(d->d_conv); cudaFree(d->d_counters);
    cudaFree(d->res_work); cudaFree(d->res_conv_buf);
    for (int k = 0; k < 2; k++) {
        ldpc_b200_decoder::PipeSet &s = d->pipe[k];
        if (k == 1) { cudaFree(s.d_in); cudaFree(s.d_bits); cudaFree(s.d_iters); cudaFree(s.d_conv); }
        cudaFree(s.d_bytes);
        if (s.h_conv) cudaFreeHost(s.h_conv);
        if (s.h2d_done) { cudaEventDestroy(s.h2d_done); cudaEventDestroy(s.d2h_done); cudaEventDestroy(s.t0); cudaEventDestroy(s.t1); }
    }
    if (d->d2h_stream) cudaStreamDestroy(d->d2h_stream);
    if (d->h_pinned) cudaFreeHost(d->h_pinned);
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    {   // the code may have lost its last handle while this decoder was alive: then it goes now
        ldpc_b200_code *c = const_cast<ldpc_b200_code *>(d->code);
        bool orphan = false;
        { std::lock_guard<std::mutex> lk(g_reg_mu); if (g_live_codes.count(c)) { c->dec_refs--; orphan = c->refs == 0 && c->dec_refs == 0; } }
        if (orphan) ldpc_b200_code_destroy(c);
    }
    delete d;
}

void *ldpc_b200_decoder_stream(ldpc_b200_decoder_t *d) { return d ? (void *)d->stream : nullptr; }

int ldpc_b200_decoder_sync(ldpc_b200_decoder_t *d)
{
    if (!d) { set_error("decoder_sync: null decoder"); return LDPC_B200_EINVAL; }
    LDPC_CUDA(cudaStreamSynchronize(d->stream));
    return finish_timing(d);
}

int ldpc_b200_decoder_mailbox_stats(const ldpc_b200_decoder_t *d, uint64_t *launches, uint64_t *frames)
{
    if (!d) { set_error("mailbox_stats: null decoder"); return LDPC_B200_EINVAL; }
    mailbox_stats(d, launches, frames);
    return 0;
}

int ldpc_b200_decoder_last_timing(ldpc_b200_decoder_t *d, double *kernel_ms, int64_t *launches)
{
    if (!d) { set_error("last_timing: null decoder"); return LDPC_B200_EINVAL; }
    int rc = finish_timing(d);
    if (rc) return rc;
    if (kernel_ms) *kernel_ms = d->last_ms;
    if (launches) *launches = d->last_launches;
    return 0;
}

int ldpc_b200_decode_device(ldpc_b200_decoder_t *d, int max_iters, int early_exit, int n_frames, const void *d_llr,
                            uint32_t *d_bits_packed, int32_t *d_iters, uint8_t *d_converged)
{
    if (!d || max_iters < 0 || n_frames < 0 || (n_frames && (!d_llr || !d_bits_packed))) { set_error("decode_device: bad arguments"); return LDPC_B200_EINVAL; }
    if (n_frames > d->max_batch) { set_error("decode_device: n_frames %d > max_batch %d", n_frames, d->max_batch); return LDPC_B200_EINVAL; }
    if (n_frames == 0) return 0;
    LDPC_CUDA(cudaSetDevice(d->code->device));
    int rc = ensure_staging(d);
    if (rc) return rc;
    begin_timing(d);
    rc = decode_dispatch(d, max_iters, early_exit, n_frames, d_llr, d->precision == LDPC_B200_PREC_FP64, d_bits_packed,
                         d_iters ? d_iters : d->d_iters, d_converged ? d_converged : d->d_conv);
    end_timing(d);
    return rc;
}

// ---- host-side narrowing of float64 LLRs (fp32 decoders) -------------------------------------------------------
// The reference hands the decoder `U.Vector Double` (Utils.hs:68); the production kernels compute in float32.  The
// host->device hop is the bottleneck of the host-buffer calls (PCIe, ~50 GB/s), so the doubles are rounded to float
// (round-to-nearest, the same conversion the device would do) by a small pool of host threads into pinned staging
// and only 4 bytes per LLR cross PCIe.  OPT-IN (LDPC_B200_HOST_NARROW=1): on the 16-vCPU host of the B200 boxes the
// narrowing itself runs at ~36 GB/s of input and the call got SLOWER (3.6 vs 4.9 Gbit/s for jpl.4096.4.5, 8 threads:
// 4.0), so the default keeps the doubles on PCIe and converts on the device; hosts with more memory bandwidth per GPU
// can turn it on.
class HostPool {
public:
    static HostPool &get() { static HostPool p; return p; }
    int size() const { return (int)workers_.size(); }
    // runs fn(part, nparts) on every worker and returns when all are done
    void run(const std::function<void(int, int)> &fn)
    {
        std::lock_guard<std::mutex> one_job(run_mu_);      // done_.wait releases mu_: callers are serialised here instead
        std::unique_lock<std::mutex> lk(mu_);
        fn_ = &fn; pending_ = (int)workers_.size(); gen_++;
        cv_.notify_all();
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }
private:
    HostPool()
    {
        unsigned hc = std::thread::hardware_concurrency();
        int n = (int)std::min<unsigned>(hc ? hc : 4, 16);
        if (const char *e = getenv("LDPC_B200_HOST_THREADS")) n = std::max(1, atoi(e));
        for (int i = 0; i < n; i++) workers_.emplace_back([this, i, n] { loop(i, n); });
    }
    ~HostPool()
    {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void loop(int i, int n)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)> *fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            if (fn) (*fn)(i, n);
            { std::lock_guard<std::mutex> lk(mu_); if (--pending_ == 0) done_.notify_all(); }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)> *fn_ = nullptr;
    int pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

static bool host_narrow_enabled()
{
    static int on = -1;
    if (on < 0) { const char *e = getenv("LDPC_B200_HOST_NARROW"); on = (e && e[0] == '1') ? 1 : 0; }
    return on == 1;
}

// dst[f][0..take) = (float)src[f*stride + 0..take), rows of `take` floats packed back to back
static void narrow_rows(const double *src, size_t stride, int take, int F, float *dst)
{
    HostPool::get().run([&](int part, int nparts) {
        int f0 = (int)((int64_t)F * part / nparts), f1 = (int)((int64_t)F * (part + 1) / nparts);
        for (int f = f0; f < f1; f++) {
            const double *in = src + (size_t)f * stride;
            float *out = dst + (size_t)f * take;
            for (int j = 0; j < take; j++) out[j] = (float)in[j];
        }
    });
}

static const int PIPE_CHUNK_MAX = 4096;

static int ensure_pipe(ldpc_b200_decoder *d, bool want_bytes)
{
    int rc = ensure_staging(d);
    if (rc) return rc;
    int N = d->code->N;
    if (!d->d2h_stream) LDPC_CUDA(cudaStreamCreateWithFlags(&d->d2h_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
        ldpc_b200_decoder::PipeSet &s = d->pipe[k];
        if (!s.h2d_done) {
            LDPC_CUDA(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
            LDPC_CUDA(cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming));
            LDPC_CUDA(cudaEventCreate(&s.t0));
            LDPC_CUDA(cudaEventCreate(&s.t1));
            s.frames = std::min(d->max_batch, PIPE_CHUNK_MAX);
            if (k == 0) {              // set 0 aliases the decoder's own staging (sized for max_batch)
                s.d_in = d->d_in; s.d_bits = d->d_bits; s.d_iters = d->d_iters; s.d_conv = d->d_conv;
            } else {
                LDPC_CUDA(cudaMalloc(&s.d_in, (size_t)s.frames * N * sizeof(double)));
                LDPC_CUDA(cudaMalloc(&s.d_bits, (size_t)s.frames * words_per_frame(N) * sizeof(uint32_t)));
                LDPC_CUDA(cudaMalloc(&s.d_iters, (size_t)s.frames * sizeof(int32_t)));
                LDPC_CUDA(cudaMalloc(&s.d_conv, (size_t)s.frames));
            }
        }
        if (want_bytes && !s.d_bytes) LDPC_CUDA(cudaMalloc(&s.d_bytes, (size_t)s.frames * N));
    }
    return 0;
}

// Shared body of the two host-buffer entry points: frames go through in chunks on three streams so that the
// H2D copy of chunk k+1, the decode of chunk k and the D2H copy of chunk k-1 overlap (pinned host buffers make
// the copies truly asynchronous; pageable ones still work).
// in_stride / in_take (elements): host row stride and number of leading LLRs per frame that exist on the host; the
// remaining N - in_take positions are punctured (LLR 0, Utils.hs:55) and never cross PCIe.  out_take: decision bytes
// per frame copied back (the first m_length, Utils.hs:72), host row stride = out_take.  0 = whole rows (N).
static int decode_host(ldpc_b200_decoder_t *d, int max_iters, int n_frames, const void *llr, bool is_double,
                       uint8_t *bits_bytes, uint32_t *bits_packed, int32_t *iters, uint8_t *converged,
                       int in_stride = 0, int in_take = 0, int out_take = 0)
{
    if (!d || max_iters < 0 || n_frames < 0 || (n_frames && (!llr || (!bits_bytes && !bits_packed)))) { set_error("decode: bad arguments"); return LDPC_B200_EINVAL; }
    if (n_frames == 0) return 0;
    if (is_double && n_frames <= MAILBOX_DIRECT_FRAMES && mailbox_enabled() && pick_engine(d) == LDPC_B200_ENGINE_RESIDENT) {
        // the tester's pattern (one frame per call, Utils.hs:68-69): no copies, no stream synchronisation -- the frame goes
        // through the mailbox its decoder handle shares with the other handles of the code (mailbox.cu)
        const int Nn = d->code->N;
        int rc = mailbox_decode(d, d->engine, max_iters, n_frames, (const double *)llr, in_take > 0 ? in_stride : Nn, in_take > 0 ? in_take : Nn,
                                bits_bytes, out_take, bits_packed, iters, converged);
        d->last_ms = 0.0; d->last_launches = 1; d->timing_pending = false;
        return rc;
    }
    LDPC_CUDA(cudaSetDevice(d->code->device));
    int rc = ensure_pipe(d, bits_bytes != nullptr);
    if (rc) return rc;
    int N = d->code->N, W = words_per_frame(N);
    size_t in_es = is_double ? sizeof(double) : sizeof(float);
    int cap = d->pipe[0].frames;
    static const int pipe_div = [] { const char *e = getenv("LDPC_B200_PIPE_DIV"); return e ? std::max(1, atoi(e)) : 16; }();   // (thread-safe magic static) 16 chunks: the un-overlapped tail (decode + D2H of the last chunk) is small; measured 4: 4.92, 8: 5.12, 16: 5.14, 32: 5.00 Gbit/s
    int chunk = std::max(256, (n_frames + pipe_div - 1) / pipe_div);
    chunk = std::min(chunk, cap);
    // float64 in, float32 decoder: round on the host, ship half the bytes (large batches only: the pool has a wake-up cost)
    const bool narrow = is_double && d->precision == LDPC_B200_PREC_FP32 && n_frames >= 64 && host_narrow_enabled();
    const bool packed_rows = !narrow && is_double && d->precision == LDPC_B200_PREC_FP32 && in_take > 0 && in_take < N &&
                             in_stride == in_take && pick_engine(d) == LDPC_B200_ENGINE_RESIDENT;
    double total_ms = 0.0;
    int64_t total_launches = 0;
    auto harvest = [&](ldpc_b200_decoder::PipeSet &s) -> int {
        if (!s.timing_pending) return 0;
        LDPC_CUDA(cudaEventSynchronize(s.d2h_done));
        float ms = 0.f;
        LDPC_CUDA(cudaEventElapsedTime(&ms, s.t0, s.t1));
        total_ms += ms;
        s.timing_pending = false;
        return 0;
    };
    int k = 0;
    for (int f0 = 0; f0 < n_frames && !rc; f0 += chunk, k++) {
        int F = std::min(chunk, n_frames - f0);
        ldpc_b200_decoder::PipeSet &s = d->pipe[k & 1];
        rc = harvest(s);                         // also guarantees the set's buffers are free again
        if (rc) break;
        bool dev_double = is_double;
        int pk_stride = 0, pk_take = 0;
        if (packed_rows) {
            // contiguous rows of in_take doubles: ONE 1-D copy; the device conversion pass spreads them to rows of N floats
            // and writes the punctured zeros
            LDPC_CUDA(cudaMemcpyAsync(s.d_in, (const char *)llr + (size_t)f0 * in_take * in_es, (size_t)F * in_take * in_es, cudaMemcpyHostToDevice, d->copy_stream));
            pk_stride = pk_take = in_take;
        } else if (narrow) {
            // chunk k is narrowed while chunk k-1 decodes and its H2D copy runs (the calls below are asynchronous)
            if (!s.h_conv) LDPC_CUDA(cudaHostAlloc((void **)&s.h_conv, (size_t)s.frames * N * sizeof(float), cudaHostAllocDefault));
            const int take = in_take > 0 ? in_take : N;
            const size_t stride = in_take > 0 ? (size_t)in_stride : (size_t)N;
            narrow_rows((const double *)llr + (size_t)f0 * stride, stride, take, F, s.h_conv);
            if (take < N) LDPC_CUDA(cudaMemset2DAsync((char *)s.d_in + (size_t)take * 4, (size_t)N * 4, 0, (size_t)(N - take) * 4, F, d->copy_stream));
            LDPC_CUDA(cudaMemcpy2DAsync(s.d_in, (size_t)N * 4, s.h_conv, (size_t)take * 4, (size_t)take * 4, F, cudaMemcpyHostToDevice, d->copy_stream));
            dev_double = false;
        } else if (in_take > 0 && (in_take < N || in_stride != N)) {
            if (in_take < N)
                LDPC_CUDA(cudaMemset2DAsync((char *)s.d_in + (size_t)in_take * in_es, (size_t)N * in_es, 0, (size_t)(N - in_take) * in_es, F, d->copy_stream));
            LDPC_CUDA(cudaMemcpy2DAsync(s.d_in, (size_t)N * in_es, (const char *)llr + (size_t)f0 * in_stride * in_es, (size_t)in_stride * in_es,
                                        (size_t)in_take * in_es, F, cudaMemcpyHostToDevice, d->copy_stream));
        } else {
            LDPC_CUDA(cudaMemcpyAsync(s.d_in, (const char *)llr + (size_t)f0 * N * in_es, (size_t)F * N * in_es, cudaMemcpyHostToDevice, d->copy_stream));
        }
        LDPC_CUDA(cudaEventRecord(s.h2d_done, d->copy_stream));
        LDPC_CUDA(cudaStreamWaitEvent(d->stream, s.h2d_done, 0));
        d->launches = 0;
        LDPC_CUDA(cudaEventRecord(s.t0, d->stream));
        rc = decode_dispatch(d, max_iters, 1, F, s.d_in, dev_double, s.d_bits, s.d_iters, s.d_conv, pk_stride, pk_take);
        if (!rc && bits_bytes) rc = unpack_bits(d, F, s.d_bits, s.d_bytes);
        if (rc) break;
        LDPC_CUDA(cudaEventRecord(s.t1, d->stream));
        total_launches += d->launches;
        LDPC_CUDA(cudaStreamWaitEvent(d->d2h_stream, s.t1, 0));
        if (bits_bytes && out_take > 0 && out_take < N)
            LDPC_CUDA(cudaMemcpy2DAsync(bits_bytes + (size_t)f0 * out_take, out_take, s.d_bytes, N, out_take, F, cudaMemcpyDeviceToHost, d->d2h_stream));
        else if (bits_bytes) LDPC_CUDA(cudaMemcpyAsync(bits_bytes + (size_t)f0 * N, s.d_bytes, (size_t)F * N, cudaMemcpyDeviceToHost, d->d2h_stream));
        if (bits_packed) LDPC_CUDA(cudaMemcpyAsync(bits_packed + (size_t)f0 * W, s.d_bits, (size_t)F * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->d2h_stream));
        if (iters) LDPC_CUDA(cudaMemcpyAsync(iters + f0, s.d_iters, (size_t)F * sizeof(int32_t), cudaMemcpyDeviceToHost, d->d2h_stream));
        if (converged) LDPC_CUDA(cudaMemcpyAsync(converged + f0, s.d_conv, (size_t)F, cudaMemcpyDeviceToHost, d->d2h_stream));
        LDPC_CUDA(cudaEventRecord(s.d2h_done, d->d2h_stream));
        s.timing_pending = true;
    }
    for (int q = 0; q < 2; q++) {
        int r2 = harvest(d->pipe[q]);
        if (!rc) rc = r2;
    }
    if (rc) { cudaStreamSynchronize(d->copy_stream); cudaStreamSynchronize(d->stream); cudaStreamSynchronize(d->d2h_stream); }
    d->last_ms = total_ms;
    d->last_launches = total_launches;
    d->timing_pending = false;
    return rc;
}

int ldpc_b200_decode(ldpc_b200_decoder_t *d, int max_iters, int n_frames, const double *llr, uint8_t *bits,
                     int32_t *iters, uint8_t *converged)
{
    return decode_host(d, max_iters, n_frames, llr, true, bits, nullptr, iters, converged);
}

/* internal (host_ecc.cu): the mkLDPC decode closure in one pass -- only the transmitted LLRs go up, only the message
 * bits come back (Utils.hs:55,68-72). */
int ldpc_b200_decode_punctured(ldpc_b200_decoder_t *d, int max_iters, int n_frames, const double *rx, int rx_stride, int take,
                               uint8_t *msg_bits, int k_take)
{
    if (!d) { set_error("decode_punctured: bad arguments"); return LDPC_B200_EINVAL; }
    const int N = d->code->N;
    if (take < 1 || take > N || rx_stride < take || k_take < 1 || k_take > N) {
        set_error("decode_punctured: need 1 <= take <= N = %d, rx_stride >= take, 1 <= k_take <= N (got take %d, stride %d, k_take %d)", N, take, rx_stride, k_take);
        return LDPC_B200_EINVAL;
    }
    return decode_host(d, max_iters, n_frames, rx, true, msg_bits, nullptr, nullptr, nullptr, rx_stride, take, k_take);
}

int ldpc_b200_decode_f32(ldpc_b200_decoder_t *d, int max_iters, int n_frames, const float *llr, uint32_t *bits_packed,
                         int32_t *iters, uint8_t *converged)
{
    return decode_host(d, max_iters, n_frames, llr, false, nullptr, bits_packed, iters, converged);
}

int ldpc_b200_decode_trace(ldpc_b200_decoder_t *d, int max_iters, const double *llr, double *c2v, double *lam,
                           int32_t *iters, uint8_t *bits, uint8_t *converged)
{
    if (!d || max_iters < 0 || !llr) { set_error("decode_trace: bad arguments"); return LDPC_B200_EINVAL; }
    LDPC_CUDA(cudaSetDevice(d->code->device));
    int rc = ensure_staging(d);
    if (rc) return rc;
    int N = d->code->N, W = words_per_frame(N);
    if (pick_engine(d) == LDPC_B200_ENGINE_RESIDENT) {
        // the tap of the decoder's OWN kernel: TRACE instance of the compiled RESIDENT specialisation (resident_trace.cu) or
        // the run-time flag of the ONCHIP kernel; messages come back in the decoder's precision and are widened here
        rc = trace_onchip_engines(d, max_iters, llr, c2v, lam);
        if (rc) return rc;
    } else {
        LDPC_CUDA(cudaMemcpyAsync(d->d_in, llr, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, d->stream));
        begin_timing(d);
        rc = stream_decode(d, max_iters, 1, 1, d->d_in, true, d->d_bits, d->d_iters, d->d_conv, c2v, lam);
        end_timing(d);
        if (rc) return rc;
    }
    std::vector<uint32_t> packed(W);
    int32_t it = 0; uint8_t cv = 0;
    LDPC_CUDA(cudaMemcpyAsync(packed.data(), d->d_bits, W * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->stream));
    LDPC_CUDA(cudaMemcpyAsync(&it, d->d_iters, sizeof it, cudaMemcpyDeviceToHost, d->stream));
    LDPC_CUDA(cudaMemcpyAsync(&cv, d->d_conv, 1, cudaMemcpyDeviceToHost, d->stream));
    LDPC_CUDA(cudaStreamSynchronize(d->stream));
    finish_timing(d);
    if (bits) for (int j = 0; j < N; j++) bits[j] = (packed[j >> 5] >> (j & 31)) & 1u;
    if (iters) *iters = it;
    if (converged) *converged = cv;
    return 0;
}

int ldpc_b200_generate_llr(ldpc_b200_decoder_t *d, double ebn0_db, int k_info, int n_tx, uint64_t seed,
                           uint64_t first_frame, int n_frames, const uint32_t *d_codewords, void *d_llr)
{
    if (!d || !d_llr || n_frames < 0 || k_info <= 0 || n_tx <= 0 || n_tx > d->code->N) { set_error("generate_llr: bad arguments"); return LDPC_B200_EINVAL; }
    LDPC_CUDA(cudaSetDevice(d->code->device));
    float sigma, scale;
    channel_params(ebn0_db, k_info, n_tx, &sigma, &scale);
    return channel_generate(d, sigma, scale, n_tx, seed, first_frame, n_frames, d_codewords, d_llr);
}

int ldpc_b200_simulate_device(ldpc_b200_decoder_t *d, int max_iters, int early_exit, double ebn0_db, int k_info, int n_tx,
                              uint64_t seed, uint64_t first_frame, uint64_t n_frames, uint64_t *d_counters)
{
    if (!d || !d_counters || max_iters < 0 || k_info <= 0 || k_info > d->code->N || n_tx <= 0 || n_tx > d->code->N) { set_error("simulate: bad arguments"); return LDPC_B200_EINVAL; }
    LDPC_CUDA(cudaSetDevice(d->code->device));
    int rc = ensure_staging(d);
    if (rc) return rc;
    float sigma, scale;
    channel_params(ebn0_db, k_info, n_tx, &sigma, &scale);
    begin_timing(d);
    if (pick_engine(d) == LDPC_B200_ENGINE_RESIDENT) {
        rc = resident_simulate(d, max_iters, early_exit, sigma, scale, k_info, n_tx, seed, first_frame, n_frames, d_counters);
    } else {
        // generate -> decode -> count, one max_batch at a time; frames are generated in the decoder's precision
        // straight into the staging buffer.
        for (uint64_t f0 = 0; f0 < n_frames && !rc; f0 += (uint64_t)d->max_batch) {
            int F = (int)std::min<uint64_t>((uint64_t)d->max_batch, n_frames - f0);
            rc = channel_generate(d, sigma, scale, n_tx, seed, first_frame + f0, F, nullptr, d->d_in);
            if (!rc) rc = stream_decode(d, max_iters, early_exit, F, d->d_in, d->precision == LDPC_B200_PREC_FP64, d->d_bits, d->d_iters, d->d_conv, nullptr, nullptr);
            if (!rc) rc = channel_count(d, F, k_info, d->d_bits, nullptr, d->d_iters, d->d_conv, d_counters);
        }
    }
    end_timing(d);
    return rc;
}

int ldpc_b200_simulate(ldpc_b200_decoder_t *d, int max_iters, int early_exit, double ebn0_db, int k_info, int n_tx,
                       uint64_t seed, uint64_t first_frame, uint64_t n_frames, ldpc_b200_counters_t *out)
{
    if (!d || !out) { set_error("simulate: bad arguments"); return LDPC_B200_EINVAL; }
    LDPC_CUDA(cudaSetDevice(d->code->device));
    int rc = ensure_staging(d);
    if (rc) return rc;
    LDPC_CUDA(cudaMemsetAsync(d->d_counters, 0, 8 * sizeof(uint64_t), d->stream));
    rc = ldpc_b200_simulate_device(d, max_iters, early_exit, ebn0_db, k_info, n_tx, seed, first_frame, n_frames, d->d_counters);
    if (rc) return rc;
    uint64_t h[8];
    LDPC_CUDA(cudaMemcpyAsync(h, d->d_counters, sizeof h, cudaMemcpyDeviceToHost, d->stream));
    LDPC_CUDA(cudaStreamSynchronize(d->stream));
    finish_timing(d);
    out->frames += h[0]; out->frame_errors += h[1]; out->bit_errors += h[2]; out->sum_iters += h[3]; out->unconverged += h[4];
    out->kernel_ms += d->last_ms;
    return 0;
}

int ldpc_b200_count_errors(ldpc_b200_decoder_t *d, int n_frames, int k_info, const uint32_t *d_bits_packed,
                           const uint32_t *d_ref_packed, const int32_t *d_iters, const uint8_t *d_converged, uint64_t *d_counters)
{
    if (!d || n_frames < 0 || k_info < 0 || k_info > d->code->N || !d_bits_packed || !d_iters || !d_converged || !d_counters) { set_error("count_errors: bad arguments"); return LDPC_B200_EINVAL; }
    LDPC_CUDA(cudaSetDevice(d->code->device));
    return channel_count(d, n_frames, k_info, d_bits_packed, d_ref_packed, d_iters, d_converged, d_counters);
}

int ldpc_b200_malloc(void **p, uint64_t bytes)
{
    if (!p) return LDPC_B200_EINVAL;
    int rc = require_init();
    if (rc) return rc;
    LDPC_CUDA(cudaMalloc(p, bytes ? bytes : 1));
    return 0;
}
int ldpc_b200_free(void *p) { LDPC_CUDA(cudaFree(p)); return 0; }
int ldpc_b200_malloc_host(void **p, uint64_t bytes)
{
    if (!p) return LDPC_B200_EINVAL;
    int rc = require_init();
    if (rc) return rc;
    LDPC_CUDA(cudaMallocHost(p, bytes ? bytes : 1));
    return 0;
}
int ldpc_b200_free_host(void *p) { LDPC_CUDA(cudaFreeHost(p)); return 0; }
int ldpc_b200_memcpy_h2d(void *dst, const void *src, uint64_t bytes) { LDPC_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)); return 0; }
int ldpc_b200_memcpy_d2h(void *dst, const void *src, uint64_t bytes) { LDPC_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); return 0; }
int ldpc_b200_memset(void *p, int value, uint64_t bytes) { LDPC_CUDA(cudaMemset(p, value, bytes)); return 0; }

int ldpc_b200_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    if (!ctr || !key || !out) return LDPC_B200_EINVAL;
    Philox4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    for (int i = 0; i < 4; i++) out[i] = r.v[i];
    return 0;
}

}  // extern "C"
