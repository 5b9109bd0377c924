// resident_trace.cu -- message-tap instances of the RESIDENT kernels (TRACE = true): the same decode, with the
// check->variable messages ne' (in reference edge order: row-major, ascending column, Orig.hs:82-92) and the posteriors
// lam' of every executed iteration of ONE frame written to global memory.  Separate translation unit so that the
// production instances of resident_engine.cu carry none of it (template bool: zero cost when off).
// Used by ldpc_b200_decode_trace for decoders whose engine is RESIDENT: the float32 1e-4 / float64 1e-9 per-edge bars of
// BASELINE.json are checked on the kernels that produce the throughput numbers.
#include "common.cuh"
#include "philox.cuh"
#include <atomic>
#include <utility>
#include <type_traits>

#ifndef RES_SPA_LOOPED
#define RES_SPA_LOOPED 1
#endif
#ifndef RES_FOLD
#define RES_FOLD 1
#endif

namespace ldpc {

#include "qc_tables.inc"
#include "bp_tree.cuh"
#include "resident_common.cuh"
#include "tma.cuh"
#include "resident_kernel.cuh"

template <typename T, int ALGO>
static int trace_kind(ldpc_b200_decoder *d, ResidentParams &p)
{
    switch (d->code->resident_kind) {
    case 1: return launch_instance<typename PlanFor<QcJpl1024P2, T, ALGO>::type, T, 4, ALGO, sizeof(T) == 4 ? 2 : 1, true>(d, p, d->stream, true);
    case 2: return launch_instance<typename PlanFor<QcJpl4096P2, T, ALGO>::type, T, 1, ALGO, sizeof(T) == 4 ? 2 : 1, true>(d, p, d->stream, true);
    default: set_error("no RESIDENT specialisation for this code"); return LDPC_B200_EUNSUPPORTED;
    }
}

// One frame (d_llr: N values in the decoder's precision) through the TRACE instance of the decoder's own kernel.
// d_c2v: [max_iters][E], d_lam: [max_iters][N] in the decoder's precision.
int resident_trace(ldpc_b200_decoder *d, int max_iters, const void *d_llr, uint32_t *d_bits, int32_t *d_iters, uint8_t *d_conv,
                   void *d_c2v, void *d_lam)
{
    ResidentParams p = {};
    p.llr = d_llr; p.llr_bulk = 0; p.bits = d_bits; p.iters = d_iters; p.conv = d_conv;
    p.n_frames = 1; p.max_iters = max_iters; p.early_exit = 1; p.k_info = 0; p.n_tx = d->code->N;
    p.trace_c2v = d_c2v; p.trace_lam = d_lam; p.trace_ne = d->code->E;
    if (d->algo == LDPC_B200_ALGO_SPA) return trace_kind<float, LDPC_B200_ALGO_SPA>(d, p);
    return d->precision == LDPC_B200_PREC_FP64 ? trace_kind<double, LDPC_B200_ALGO_MINSUM>(d, p)
                                               : trace_kind<float, LDPC_B200_ALGO_MINSUM>(d, p);
}

}  // namespace ldpc
