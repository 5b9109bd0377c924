// resident_engine.cu -- RESIDENT engine: flooding min-sum for the quasi-cyclic JPL AR4JA codes with the
// whole frame state on chip for all iterations.
//
// Why: the reference's state model moves 3E+3N words per frame and iteration; streamed from
// HBM that caps a B200 at ~1.06 M frames/s for jpl.4096.4.5 at 20 iterations.  Here a frame costs HBM only
// its LLRs in and its packed decisions out; the binding resources are instruction issue, the ALU pipe and
// shared-memory bandwidth.
//
// Mapping (SZ = circulant size, NP = threads per check position, one frame = NP*SZ threads):
//   * thread (part, i): check position i of the block rows of its part (tables in qc_tables.inc; the split gives
//     every part the same number of circulants).  Its check->variable messages ne[] live in REGISTERS for the
//     whole decode (one register per edge, compile-time indexed after full unrolling).
//   * the negated posterior mu = -lam and negL = -L live in SHARED memory.  Circulant (r,k) with shift s connects
//     check (r,i) to variable (c, (i+s) mod SZ): a warp reads/writes 32 consecutive words -> conflict-free.
//     Every block column is a SZ*sizeof(T)-aligned block, so the rotated address is formed with one add and one
//     bit-select (LOP3); shifts are compile-time constants and the 57 shift-0 circulants cost nothing.
//   * check phase (Min.hs:75-89): x = ne + mu gathered from smem; syndrome parity = xor of the sign bits of mu
//     (Orig.hs:73-78); leave-one-out reduction with the reference's min' operator (one FMNMX.XORSIGN each) in a
//     binary tree; ne' = -3/4 * (reduction of the others), in place.
//   * vote: barrier.red.or over the frame's threads = "syndrome all zero?"  (Orig.hs:68-69).
//   * variable phase (Orig.hs:95-98): the right fold ne'[m1] + (ne'[m2] + ... (ne'[mk] + L_j)) is built in smem with the
//     reference's operands, order and roundings by a PLAN made offline (tools/gen_qc_tables.py; VnPlan in
//     resident_kernel.cuh): at most one chain link per column (load the running value, apply the own message, store);
//     every other non-final edge is a singleton -- its thread stores the message into a slot of its own right behind the
//     row's tree in the CHECK phase, where the store overlaps ALU-bound work --; the column's executor thread loads the
//     chain value (or negL) and the column's slots and folds them in chain order, and folds edges of the same variable
//     that it owns itself (same part, same shift) straight from registers.  2 stages + the vote = 3 barriers per
//     iteration (the original one-link-per-stage schedule: 7).  Within a stage every address is touched by exactly one
//     thread; a stage issues all its loads before its first store (the accesses are ordered volatile asm).  Same
//     association as the reference, so the fp64 instance reproduces the reference bit for bit under every plan.
//   * last-link fold (min-sum): the executor finishes the new posterior of its variable and owns one (or more) of its
//     edges, so it keeps x = ne' + mu and the sign of mu for those edges in registers: 56 of the 156 check-phase loads of
//     an iteration never happen.
//   * early exit per frame; frames are handed out by an atomic work counter (persistent CTAs); host-supplied frames:
//     the NEXT frame's LLR row is fetched by one bulk copy (cp.async.bulk + mbarrier) while the current one decodes.
//   * sum-product (float32) runs the same phases in the complement domain, with one shared copy of the degree-18 row
//     arithmetic (cn_phase_spa_looped) because the fully unrolled body did not fit the instruction cache.
//
// fp32 is the production instance; fp64 exists to prove the index logic bit-exactly.
#include "common.cuh"
#include "philox.cuh"
#include <atomic>
#include <utility>
#include <type_traits>

// RES_PARTS: threads per check position (2 or 4). More parts = fewer message registers per thread = more resident
// warps to hide shared-memory latency, at the price of more barrier participants.
#ifndef RES_PARTS
#define RES_PARTS 2
#endif
// RES_SPA_LOOPED: sum-product check phase with one shared copy of the big-row arithmetic (see cn_phase_spa_looped)
#ifndef RES_SPA_LOOPED
#define RES_SPA_LOOPED 1
#endif
// RES_FOLD: min-sum check phase takes x of the last-link edges from the variable phase (see cn_phase)
#ifndef RES_FOLD
#define RES_FOLD 1
#endif

namespace ldpc {

#include "qc_tables.inc"
#include "bp_tree.cuh"
#include "resident_common.cuh"
#include "tma.cuh"

#include "resident_kernel.cuh"

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

#if RES_PARTS == 2
typedef QcJpl1024P2 QcJpl1024;
typedef QcJpl4096P2 QcJpl4096;
#else
typedef QcJpl1024P4 QcJpl1024;
typedef QcJpl4096P4 QcJpl4096;
#endif

template <class C>
static bool matches(const ldpc_b200_code *c)
{
    if (c->sz != C::SZ || c->brows != C::BR || c->bcols != C::BC) return false;
    for (int r = 0; r < C::BR; r++) {
        int k = 0;
        for (int b = 0; b < C::BC; b++) {
            int s = c->shifts[r * C::BC + b];
            if (s < 0) continue;
            if (k >= C::row_deg[r] || C::row_col[r][k] != b || C::row_shift[r][k] != s) return false;
            k++;
        }
        if (k != C::row_deg[r]) return false;
    }
    return true;
}

int resident_kind_for(const ldpc_b200_code *c)
{
    if (c->sz == 0) return 0;
    if (matches<QcJpl1024>(c)) return 1;
    if (matches<QcJpl4096>(c)) return 2;
    return 0;
}

// which on-chip kernel serves this decoder: the compiled specialisation of its code, or (no specialisation, or the
// ONCHIP engine asked for explicitly) the run-time table kernel of onchip_engine.cu
int resident_kind_of(const ldpc_b200_decoder *d)
{
    return d->engine == LDPC_B200_ENGINE_ONCHIP ? 3 : d->code->resident_kind;
}

bool resident_supports(const ldpc_b200_decoder *d)
{
    // min-sum: fp32 and fp64; sum-product: fp32 only (fp64 sum-product is the STREAM engine's literal form)
    if (d->algo != LDPC_B200_ALGO_SPA && d->algo != LDPC_B200_ALGO_MINSUM) return false;   // the arraylet-order statements live on STREAM
    if (!(d->algo == LDPC_B200_ALGO_MINSUM || d->precision == LDPC_B200_PREC_FP32)) return false;
    if (resident_kind_of(d) == 3) return onchip_fits(d->code, d->precision);
    return d->code->resident_kind != 0;
}

// float64 rows of `take` values at a stride of `stride` -> float rows of N values, zero beyond `take` (the punctured
// positions, Utils.hs:55)
template <typename T>
__global__ void k_convert_f64(const double *__restrict__ in, T *__restrict__ out, size_t n, int N, int stride, int take)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const size_t f = idx / (size_t)N;
    const int j = (int)(idx - f * (size_t)N);
    out[idx] = j < take ? (T)in[f * (size_t)stride + j] : (T)0;
}

template <class C, typename T, int FPC, int ALGO>
static int launch(ldpc_b200_decoder *d, ResidentParams &p, cudaStream_t stream, bool static_grid)
{
    return launch_instance<typename PlanFor<C, T, ALGO>::type, T, FPC, ALGO, sizeof(T) == 4 ? 2 : 1, false>(d, p, stream, static_grid);
}

template <typename T, int ALGO>
static int launch_kind(ldpc_b200_decoder *d, ResidentParams &p, cudaStream_t stream, bool static_grid)
{
    switch (resident_kind_of(d)) {
    case 1: return launch<QcJpl1024, T, 4, ALGO>(d, p, stream, static_grid);
    case 2: return launch<QcJpl4096, T, 1, ALGO>(d, p, stream, static_grid);
    case 3: return onchip_launch(d, p, stream, static_grid);
    default: set_error("no RESIDENT specialisation for this code"); return LDPC_B200_EUNSUPPORTED;
    }
}

static int launch_prec(ldpc_b200_decoder *d, ResidentParams &p, cudaStream_t stream = nullptr, bool static_grid = false)
{
    if (!stream) stream = d->stream;
    if (d->algo == LDPC_B200_ALGO_SPA) return launch_kind<float, LDPC_B200_ALGO_SPA>(d, p, stream, static_grid);
    return d->precision == LDPC_B200_PREC_FP64 ? launch_kind<double, LDPC_B200_ALGO_MINSUM>(d, p, stream, static_grid)
                                               : launch_kind<float, LDPC_B200_ALGO_MINSUM>(d, p, stream, static_grid);
}

// A small batch on `stream` with a static one-pass grid (no work counter): the mailbox's launch for decoders without a
// latency instance (float64, ONCHIP codes).
int resident_static_launch(ldpc_b200_decoder *d, ResidentParams &p, cudaStream_t stream)
{
    if (!resident_supports(d)) { set_error("RESIDENT engine does not support this decoder"); return LDPC_B200_EUNSUPPORTED; }
    return launch_prec(d, p, stream, true);
}

int resident_decode(ldpc_b200_decoder *d, int max_iters, int early_exit, int n_frames, const void *d_llr,
                    bool llr_is_double, uint32_t *d_bits_packed, int32_t *d_iters, uint8_t *d_conv, int src_stride, int src_take)
{
    if (!resident_supports(d)) { set_error("RESIDENT engine does not support this decoder"); return LDPC_B200_EUNSUPPORTED; }
    int N = d->code->N;
    const void *src = d_llr;
    if (llr_is_double && d->precision == LDPC_B200_PREC_FP32) {
        size_t need = (size_t)d->max_batch * N * sizeof(float);
        if (d->res_conv_bytes < need) {
            if (d->res_conv_buf) cudaFree(d->res_conv_buf);
            d->res_conv_buf = nullptr; d->res_conv_bytes = 0;
            LDPC_CUDA(cudaMalloc(&d->res_conv_buf, need));
            d->res_conv_bytes = need;
        }
        size_t n = (size_t)n_frames * N;
        const int stride = src_take > 0 ? src_stride : N, take = src_take > 0 ? src_take : N;
        k_convert_f64<float><<<(unsigned)((n + 255) / 256), 256, 0, d->stream>>>((const double *)d_llr, (float *)d->res_conv_buf, n, N, stride, take);
        d->launches++;
        src = d->res_conv_buf;
    } else if (!llr_is_double && d->precision == LDPC_B200_PREC_FP64) {
        set_error("fp64 decoder needs float64 LLRs");
        return LDPC_B200_EINVAL;
    } else if (src_take > 0) {
        set_error("packed source rows need the float64 -> float32 conversion pass");
        return LDPC_B200_EINVAL;
    }
    ResidentParams p = {};
    
    p.llr = src; p.bits = d_bits_packed; p.iters = d_iters; p.conv = d_conv; p.counters = nullptr;
    {
        const size_t row = (size_t)N * elem_size(d->precision);
        p.llr_bulk = ((uintptr_t)src % 16 == 0 && row % 16 == 0) ? 1 : 0;     // cp.async.bulk needs 16-byte aligned rows
    }
    p.n_frames = (unsigned int)n_frames; p.max_iters = max_iters; p.early_exit = early_exit;
    p.k_info = 0; p.n_tx = N;
    return launch_prec(d, p);
}

int resident_simulate(ldpc_b200_decoder *d, int max_iters, int early_exit, float sigma, float llr_scale,
                      int k_info, int n_tx, uint64_t seed, uint64_t first_frame, uint64_t n_frames,
                      uint64_t *d_counters)
{
    if (!resident_supports(d)) { set_error("RESIDENT engine does not support this decoder"); return LDPC_B200_EUNSUPPORTED; }
    const uint64_t chunk = 1ull << 30;
    for (uint64_t f0 = 0; f0 < n_frames; f0 += chunk) {
        ResidentParams p = {};
        
        p.llr = nullptr; p.bits = nullptr; p.iters = nullptr; p.conv = nullptr;
        p.counters = (unsigned long long *)d_counters;
        p.n_frames = (unsigned int)std::min<uint64_t>(chunk, n_frames - f0);
        p.max_iters = max_iters; p.early_exit = early_exit;
        p.sigma = sigma; p.scale = llr_scale; p.n_tx = n_tx; p.k_info = k_info;
        p.seed = seed; p.first_frame = first_frame + f0;
        int rc = launch_prec(d, p);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace ldpc
