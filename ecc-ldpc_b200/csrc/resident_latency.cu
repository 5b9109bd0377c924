// resident_latency.cu -- latency instances of the RESIDENT kernels for the reference tester's calling pattern: ONE frame
// per `decode` call (Utils.hs:62-69).  A frame is spread over NP = 4 threads per check position (512 threads for
// jpl.4096.4.5, 128 for jpl.1024.4.5) with the register allocation bounded for ONE CTA per SM, so the four-part form
// keeps its messages in registers (at two CTAs per SM it spills and loses to NP = 2, which is why the throughput
// instances use NP = 2); one frame per CTA, static one-pass grid, so a handful of frames from concurrent tester threads
// occupies as many SMs.  Launched by the mailbox (mailbox.cu).
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

template <int ALGO>
static int latency_kind(ldpc_b200_decoder *d, ResidentParams &p, cudaStream_t stream)
{
    switch (d->code->resident_kind) {
    case 1: return launch_instance<typename PlanFor<QcJpl1024P4, float, ALGO, true>::type, float, 1, ALGO, 1, false>(d, p, stream, true);
    case 2: return launch_instance<typename PlanFor<QcJpl4096P4, float, ALGO, true>::type, float, 1, ALGO, 1, false>(d, p, stream, true);
    default: set_error("no RESIDENT specialisation for this code"); return LDPC_B200_EUNSUPPORTED;
    }
}

bool resident_latency_supports(const ldpc_b200_decoder *d)
{
    return (d->code->resident_kind == 1 || d->code->resident_kind == 2) && d->precision == LDPC_B200_PREC_FP32 &&
           (d->algo == LDPC_B200_ALGO_MINSUM || d->algo == LDPC_B200_ALGO_SPA);
}

// p describes a small batch (p.n_frames <= MAILBOX_MAX_BATCH); one CTA per frame on `stream`.
int resident_latency_launch(ldpc_b200_decoder *d, ResidentParams &p, cudaStream_t stream)
{
    if (!resident_latency_supports(d)) { set_error("no latency instance for this decoder"); return LDPC_B200_EUNSUPPORTED; }
    return d->algo == LDPC_B200_ALGO_SPA ? latency_kind<LDPC_B200_ALGO_SPA>(d, p, stream) : latency_kind<LDPC_B200_ALGO_MINSUM>(d, p, stream);
}

}  // namespace ldpc
