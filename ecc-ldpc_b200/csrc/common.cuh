// common.cuh -- shared host-side structures of libldpc_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/ldpc_b200.h"

namespace ldpc {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define LDPC_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t _e = (call);                                                     \
        if (_e != cudaSuccess) return ldpc::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

// Device-resident H. Edge order = reference order: row-major, ascending column
// (the order Reference/Orig.hs:82-92 visits the ones of H).
struct DeviceGraph {
    int M = 0, N = 0, E = 0, max_row_deg = 0, max_col_deg = 0;
    int32_t *row_ptr = nullptr;   // [M+1]
    int32_t *col_idx = nullptr;   // [E]  edge -> variable
    int32_t *col_ptr = nullptr;   // [N+1]
    int32_t *col_edge = nullptr;  // [E]  edges of variable j, ascending check index
    // Fixed-stride records for the STREAM kernels: one cooperative load per CTA replaces the dependent
    // row_ptr -> col_idx (col_ptr -> col_edge) chain in front of the bulk loads.
    //   row_ell [M][row_ell_d + 2] : {first edge b, degree d, col_idx[b .. b+d-1], last column repeated}; null if max_row_deg > 32
    //   col_ell [N][COL_ELL_D + 1] : {degree d, edges of the variable in DESCENDING check order (the order of the
    //                                 reference's right fold), first COL_ELL_D of them, last one repeated}
    int32_t *row_ell = nullptr, *col_ell = nullptr;
    int row_ell_d = 0;
    // ONCHIP engine: 16-bit records in the order they are staged into shared memory (layout and placement:
    // onchip_engine.cu, onchip_build_tables)
    uint16_t *oc_tab = nullptr;
    int oc_tab_elems = 0, oc_se = 0, oc_cd = 0, oc_cdr = 0;   // table length (u16), message slots per frame, staged messages per variable, entries per variable record (4 / 8)
    int32_t *oc_posvar = nullptr;                    // [N] variable at position p (tail of variables with degree > oc_cd)
    int32_t *col_slot = nullptr;                     // [E] message slot of every edge in CSC order (same tail)
    int row_regular = 0;                             // every check has exactly row_ell_d edges
    static constexpr int COL_ELL_D = 8;
};

}  // namespace ldpc

struct ldpc_b200_code {
    int M = 0, N = 0, E = 0;
    std::vector<int32_t> row_ptr, col_idx, col_ptr, col_edge;
    int max_row_deg = 0, max_col_deg = 0;
    // quasi-cyclic description (sz == 0 when the code was given as CSR)
    int sz = 0, brows = 0, bcols = 0;
    std::vector<int32_t> shifts;
    int device = -1;
    int refs = 1;                 // handles sharing this object: creating an identical code on the same device returns it again
    int dec_refs = 0;             // decoders built on it: the object outlives its last handle until they are gone too
    ldpc::DeviceGraph g;          // device copy
    int resident_kind = 0;        // 0 = none, 1 / 2 = compiled RESIDENT specialisation, 3 = run-time table ONCHIP kernel
    // ONCHIP placement, host copy (message tap: slot of reference edge e, position of variable j)
    std::vector<int32_t> oc_slot_of_edge, oc_pos_of_var;
};

namespace ldpc {

struct StreamBuffers {
    void *L = nullptr, *lam = nullptr, *ne = nullptr;   // [N][Bp], [N][Bp], [E][Bp] in T
    int32_t *iters = nullptr;    // [Bp]
    uint8_t *state = nullptr;    // [Bp] 0 = running, 1 = converged, 2 = gave up (fallback to L)
    uint8_t *unsat = nullptr;    // [Bp]
    int *active = nullptr;       // [1] frames still running
    int Bp = 0;
};

}  // namespace ldpc

struct ldpc_b200_decoder {
    const ldpc_b200_code *code = nullptr;
    int algo = 1, precision = 1, max_batch = 0, engine = 0;
    int device = -1;             // = code->device, kept here so that destroy does not depend on the code still being alive
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ldpc::StreamBuffers sb;
    // staging for host-buffer calls
    void *d_in = nullptr;        // max_batch x N (double)
    uint32_t *d_bits = nullptr;  // max_batch x ceil(N/32)
    int32_t *d_iters = nullptr;
    uint8_t *d_conv = nullptr;
    uint64_t *d_counters = nullptr;   // 5 x u64
    void *h_pinned = nullptr;    // pinned staging
    size_t h_pinned_bytes = 0;
    double last_ms = 0.0;
    int64_t last_launches = 0;
    int64_t launches = 0;        // running count within a call
    bool timing_pending = false;
    // second staging set + stream/events for the pipelined host-buffer decode (H2D | decode | D2H overlap)
    struct PipeSet {
        void *d_in = nullptr; uint32_t *d_bits = nullptr; uint8_t *d_bytes = nullptr;
        int32_t *d_iters = nullptr; uint8_t *d_conv = nullptr;
        cudaEvent_t h2d_done = nullptr, t0 = nullptr, t1 = nullptr, d2h_done = nullptr;
        float *h_conv = nullptr;     // pinned [frames][N] float: float64 LLRs narrowed on the host in front of the PCIe hop
        int frames = 0;              // capacity in frames
        bool timing_pending = false;
    } pipe[2];
    cudaStream_t d2h_stream = nullptr;
    // RESIDENT engine scratch
    unsigned int *res_work = nullptr;   // persistent-CTA work counter
    int oc_blocks_per_sm = 0, oc_n_sm = 0;   // ONCHIP: cached occupancy of this decoder's persistent grid
    void *mbox = nullptr;               // the mailbox this handle shares with the other handles of its code (mailbox.cu)
    void *res_conv_buf = nullptr;       // max_batch x N float, when float64 LLRs feed the fp32 kernel
    size_t res_conv_bytes = 0;
};

namespace ldpc {

inline int words_per_frame(int N) { return (N + 31) / 32; }
inline size_t elem_size(int precision) { return precision == LDPC_B200_PREC_FP64 ? 8 : 4; }

// ---- stream engine (stream_engine.cu) ----
int stream_alloc(ldpc_b200_decoder *d);
void stream_free(ldpc_b200_decoder *d);
// d_llr: n_frames x N in T (frame-major). Runs the whole decode; outputs are frame-major device buffers.
int stream_decode(ldpc_b200_decoder *d, int max_iters, int early_exit, int n_frames, const void *d_llr,
                  bool llr_is_double, uint32_t *d_bits_packed, int32_t *d_iters, uint8_t *d_conv,
                  double *h_trace_c2v, double *h_trace_lam);

// ---- resident engine (resident_engine.cu) ----
int resident_kind_for(const ldpc_b200_code *c);          // 0 if no specialisation matches
bool resident_supports(const ldpc_b200_decoder *d);
int resident_kind_of(const ldpc_b200_decoder *d);       // 1 / 2 compiled specialisation, 3 run-time table ONCHIP kernel
// src_take > 0 (float64 source, float32 decoder only): d_llr holds rows of src_take values at a stride of src_stride
// doubles; the remaining positions of every frame are punctured (LLR 0)
int resident_decode(ldpc_b200_decoder *d, int max_iters, int early_exit, int n_frames, const void *d_llr,
                    bool llr_is_double, uint32_t *d_bits_packed, int32_t *d_iters, uint8_t *d_conv,
                    int src_stride = 0, int src_take = 0);
int resident_simulate(ldpc_b200_decoder *d, int max_iters, int early_exit, float sigma, float llr_scale,
                      int k_info, int n_tx, uint64_t seed, uint64_t first_frame, uint64_t n_frames,
                      uint64_t *d_counters);

// message tap on the decoder's own kernel (resident_trace.cu): one frame, d_c2v [max_iters][E] / d_lam [max_iters][N] in T
int resident_trace(ldpc_b200_decoder *d, int max_iters, const void *d_llr, uint32_t *d_bits, int32_t *d_iters, uint8_t *d_conv,
                   void *d_c2v, void *d_lam);
// latency instances for small batches (resident_latency.cu)
struct ResidentParams;
bool resident_latency_supports(const ldpc_b200_decoder *d);
int resident_latency_launch(ldpc_b200_decoder *d, ResidentParams &p, cudaStream_t stream);
// small batch on `stream` with a static one-pass grid through the production instance (fp64, or no latency instance)
int resident_static_launch(ldpc_b200_decoder *d, ResidentParams &p, cudaStream_t stream);

// ---- on-chip engine for arbitrary codes (onchip_engine.cu) ----
int onchip_build_tables(ldpc_b200_code *c);
bool onchip_fits(const ldpc_b200_code *c, int precision);
int onchip_launch(ldpc_b200_decoder *d, ResidentParams &p, cudaStream_t stream = nullptr, bool static_grid = false);

// ---- mailbox: one-frame-per-call path shared by all decoder handles of a code (mailbox.cu) ----
bool mailbox_enabled();
int mailbox_decode(ldpc_b200_decoder *d, int engine, int max_iters, int n_frames, const double *llr, int stride, int take,
                   uint8_t *bits_bytes, int out_take, uint32_t *bits_packed, int32_t *iters, uint8_t *converged);
void mailbox_release(ldpc_b200_decoder *d);
void mailbox_stats(const ldpc_b200_decoder *d, uint64_t *launches, uint64_t *frames);
static const int MAILBOX_DIRECT_FRAMES = 4;      // host-buffer calls of at most this many frames take the mailbox

// ---- channel (channel.cu) ----
int channel_generate(ldpc_b200_decoder *d, float sigma, float llr_scale, int n_tx, uint64_t seed,
                     uint64_t first_frame, int n_frames, const uint32_t *d_codewords, void *d_llr);
int channel_count(ldpc_b200_decoder *d, int n_frames, int k_info, const uint32_t *d_bits_packed,
                  const uint32_t *d_ref_packed, const int32_t *d_iters, const uint8_t *d_conv, uint64_t *d_counters);
int unpack_bits(ldpc_b200_decoder *d, int n_frames, const uint32_t *d_bits_packed, uint8_t *d_bytes);

}  // namespace ldpc
