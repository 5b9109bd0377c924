// resident_common.cuh -- what the on-chip engines share (included inside namespace ldpc): launch parameters, per-CTA
// counters, named-barrier helpers.
#pragma once

#ifndef MAILBOX_MAX_BATCH
#define MAILBOX_MAX_BATCH 64
#endif

__device__ __forceinline__ void bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ bool bar_or(int id, int nthreads, bool pred)
{
    int r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %3, 0;\n\tbarrier.red.or.pred q, %1, %2, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
                 : "=r"(r) : "r"(id), "r"(nthreads), "r"((int)pred) : "memory");
    return r != 0;
}

struct ResidentParams {
    const void *llr;               // [F][N] in T, or nullptr: see llr64 / simulate
    int llr_bulk;                  // llr rows are 16-byte aligned and a multiple of 16 bytes: next frame prefetched by bulk copy
    // float64 source rows (the reference's `U.Vector Double`, Utils.hs:68), converted on the fly: frame f holds llr64_take
    // values at llr64 + f * llr64_stride; positions >= llr64_take are punctured (LLR 0, Utils.hs:55).  Used when llr is null.
    const double *llr64;
    int llr64_stride, llr64_take;
    uint32_t *bits;                // [F][N/32] or nullptr
    int32_t *iters;                // [F] or nullptr
    uint8_t *conv;                 // [F] or nullptr
    unsigned long long *counters;  // 5 x u64 or nullptr
    unsigned int *work;            // work counter (zeroed before launch); nullptr: frame = CTA index * frames per CTA + slot, one pass
    unsigned int n_frames;
    int max_iters, early_exit;
    // simulate
    float sigma, scale;
    int n_tx, k_info;
    unsigned long long seed, first_frame;
    // mailbox mode (the tester's one-frame-per-call pattern, mailbox.cu): launch frame f lives in slot slot_map[f] of
    // the arrays above (which are then in mapped host memory), and its completion is published by done[slot] = done_seq
    // after a system-scope fence.
    int use_map;
    unsigned short slot_map[MAILBOX_MAX_BATCH];
    unsigned int *done;
    unsigned int done_seq;
    // message tap (TRACE instances / ONCHIP run-time flag), frame 0 of the launch: after iteration t (0-based)
    // trace_c2v[t * trace_ne + ...] = the new check->variable messages (RESIDENT: reference edge order; ONCHIP: slot order,
    // reordered by the host), trace_lam[t * N + ...] = the new posteriors (ONCHIP: position order).  In T.
    void *trace_c2v, *trace_lam;
    int trace_ne;
};

struct FrameCounters {
    unsigned long long frames = 0, ferr = 0, berr = 0, iters = 0, unconv = 0;
};

