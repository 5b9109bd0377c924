// resident_common.cuh -- what the on-chip engines share (included inside namespace ldpc): launch parameters, per-CTA
// counters, named-barrier helpers.
#pragma once

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
    const void *llr;               // [F][N] in T, or nullptr: generate on the fly (simulate)
    int llr_bulk;                  // llr rows are 16-byte aligned and a multiple of 16 bytes: next frame prefetched by bulk copy
    uint32_t *bits;                // [F][N/32] or nullptr
    int32_t *iters;                // [F] or nullptr
    uint8_t *conv;                 // [F] or nullptr
    unsigned long long *counters;  // 5 x u64 or nullptr
    unsigned int *work;            // work counter (zeroed before launch)
    unsigned int n_frames;
    int max_iters, early_exit;
    // simulate
    float sigma, scale;
    int n_tx, k_info;
    unsigned long long seed, first_frame;
};

struct FrameCounters {
    unsigned long long frames = 0, ferr = 0, berr = 0, iters = 0, unconv = 0;
};

