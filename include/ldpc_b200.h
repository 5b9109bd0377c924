/*
 * ldpc_b200.h -- C ABI of libldpc_b200.so, the B200 (sm_100a) replacement for the decode path of
 * ku-fpg/ecc-ldpc.  Plain pointers and sizes only; no CUDA or torch types.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *
 *   ldpc_b200_init / _shutdown        the `initialize` / `finalize` pair a Code carries
 *                                     (src/ECC/Code/LDPC/Utils.hs:91-103; precedent
 *                                     src/ECC/Code/LDPC/GPU/CUDA/Arraylet2.hs:287-297)
 *   ldpc_b200_code_create_csr         `loadMatrix` result as `Matrix Bool` (Utils.hs:40,
 *                                     src/Data/BitMatrix/Loader.hs:34-39) handed to `decoder0 h`
 *   ldpc_b200_code_create_qc          `QuasiCyclic Integer` H (Loader.hs:41-46) in the circulant-offset
 *                                     form of Arraylet2.hs:325-331 (-1 = zero block)
 *   ldpc_b200_decoder_create          one `decoder0 h` call (Utils.hs:53): per-thread decoder state
 *   ldpc_b200_decode                  the `decoder` slot of mkLDPC_CodeIO (Utils.hs:95):
 *                                     Rate -> Int -> U.Vector Double -> IO (Maybe (U.Vector Bool)),
 *                                     batched over n_frames; algorithm of Reference/Orig.hs:58-98 (SPA)
 *                                     and Reference/Min.hs:54-104 (min-sum, alpha = 3/4)
 *   ldpc_b200_decode_trace            per-iteration `ne'` / `lam'` of Orig.hs:81-98 for message-level checks
 *   ldpc_b200_decode_device           same decode with buffers already resident in HBM
 *   ldpc_b200_simulate*               what ECC.Tester (ecc-manifold, external) does around `decode`:
 *                                     BPSK + AWGN frames, decode, count bit / frame errors
 *   ldpc_b200_ecc_*                   the ECC record built by mkLDPC (Utils.hs:35-75): name grammar,
 *                                     unpuncture, encode, decode, message_length, codeword_length
 *
 * Conventions (Utils.hs / Orig.hs): LLR > 0 decides bit 1; ties decide 0; the syndrome is tested before
 * every iteration including iteration 0; a frame that is still unsatisfied after max_iters iterations
 * returns the hard decision of its ORIGINAL channel LLRs (Orig.hs:68-71).
 * All functions return 0 on success and a negative LDPC_B200_E* code on failure; nothing throws or
 * aborts across this boundary.  There is no CPU fallback: without a CUDA device every call that needs
 * one fails with LDPC_B200_ENODEV.
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_B200_ABI_VERSION 1

enum {
    LDPC_B200_OK = 0,
    LDPC_B200_EINVAL = -1,    /* bad argument */
    LDPC_B200_ENODEV = -2,    /* no CUDA device / not initialised */
    LDPC_B200_ECUDA = -3,     /* CUDA runtime error, see ldpc_b200_last_error() */
    LDPC_B200_ENOMEM = -4,
    LDPC_B200_EUNSUPPORTED = -5
};

enum {
    LDPC_B200_ALGO_SPA = 0,     /* sum-product, src/ECC/Code/LDPC/Reference/Orig.hs:81-92 */
    LDPC_B200_ALGO_MINSUM = 1,  /* normalised min-sum alpha = 3/4, src/ECC/Code/LDPC/Reference/Min.hs:75-89 */
    LDPC_B200_ALGO_SPA_CM = 2,  /* sum-product with the cached StableDiv row product of `ldpc/arraylet-cm`
                                   (src/ECC/Code/LDPC/Fast/CachedMult.hs:25-56,248-264); float64 only */
    /* `ldpc/arraylet` (src/ECC/Code/LDPC/Fast/Arraylet.hs:164-186) and `ldpc/arraylet-min`
     * (Fast/ArrayletMin.hs:167-192): the check updates of SPA / MINSUM with the posterior associated as
     * L_j + (ne'[m1] + (ne'[m2] + ...)) (foldRowsMatrixlet, ArrayletMin.hs:104-108,191-192) instead of the right fold
     * of Orig.hs:95-98.  STREAM engine only (verification statements; decisions equal SPA / MINSUM up to that rounding). */
    LDPC_B200_ALGO_SPA_ARRAYLET = 3,
    LDPC_B200_ALGO_MINSUM_ARRAYLET = 4
};
enum { LDPC_B200_PREC_FP64 = 0, LDPC_B200_PREC_FP32 = 1 };
/* STREAM   : any H; messages live in HBM, frame-minor layout, one launch per phase.
 * RESIDENT : the compiled specialisations for the quasi-cyclic JPL codes; the whole frame state stays in
 *            registers / shared memory of one CTA for all iterations.
 * ONCHIP   : any H whose per-frame state (2N + M*(row class + 1) words) fits shared memory; run-time tables.
 * AUTO     : RESIDENT when a specialisation exists, else ONCHIP when the code fits, else STREAM.
 * ldpc_b200_decoder_engine reports the engine a decode will actually run on. */
enum { LDPC_B200_ENGINE_AUTO = 0, LDPC_B200_ENGINE_STREAM = 1, LDPC_B200_ENGINE_RESIDENT = 2, LDPC_B200_ENGINE_ONCHIP = 3 };

typedef struct ldpc_b200_code ldpc_b200_code_t;
typedef struct ldpc_b200_decoder ldpc_b200_decoder_t;

/* Counters of a simulation run (what the tester accumulates per Eb/N0 point). */
typedef struct ldpc_b200_counters {
    uint64_t frames;        /* frames decoded */
    uint64_t frame_errors;  /* frames with >= 1 wrong message bit */
    uint64_t bit_errors;    /* wrong bits among the first k_info bits of every frame */
    uint64_t sum_iters;     /* iterations executed, summed over frames */
    uint64_t unconverged;   /* frames that fell back to the channel hard decision */
    double kernel_ms;       /* device time of the decode kernels (CUDA events) */
} ldpc_b200_counters_t;

int ldpc_b200_abi_version(void);
const char *ldpc_b200_last_error(void);

/* device = CUDA ordinal, or -1 for the current device. Idempotent. */
int ldpc_b200_init(int device);
/* The Code's `finalize` (Utils.hs:91-103): destroys every decoder and code the host has not destroyed itself and the
 * replicas cached by ldpc_b200_simulate_multi; all handles are invalid afterwards (destroying one again is a no-op).
 * The library can be initialised again. */
void ldpc_b200_shutdown(void);
/* Number of live code / decoder handles (leak checks in tests and in the host binding's finalize). */
int ldpc_b200_live_handles(int *codes, int *decoders);
int ldpc_b200_device_count(void);

/* Codes.  Creating a code that is identical (same device, same H) to a live one returns the SAME object with one more handle
 * reference: mkLDPC builds `maxThreadCount` decoders from the same matrix (Utils.hs:53), which then share one device copy of H
 * and one mailbox.  ldpc_b200_code_destroy drops one handle reference; the object itself lives until its last handle AND its
 * last decoder are gone, so the order in which a host destroys codes and decoders does not matter.
 * H as CSR: check m has variables col_idx[row_ptr[m] .. row_ptr[m+1]) in ascending order. Arrays are copied. */
int ldpc_b200_code_create_csr(int M, int N, const int32_t *row_ptr, const int32_t *col_idx, ldpc_b200_code_t **out);
/* H as brows x bcols circulants of size sz: shifts[r*bcols+c] = k means ones at (r*sz+i, c*sz+(i+k)%sz); -1 = zero. */
int ldpc_b200_code_create_qc(int sz, int brows, int bcols, const int32_t *shifts, ldpc_b200_code_t **out);
void ldpc_b200_code_destroy(ldpc_b200_code_t *code);
int ldpc_b200_code_info(const ldpc_b200_code_t *code, int *M, int *N, int *E, int *qc_sz);

/* One decoder per host thread: owns a stream and its device buffers, so no locking is needed. */
int ldpc_b200_decoder_create(const ldpc_b200_code_t *code, int algo, int precision, int max_batch,
                             ldpc_b200_decoder_t **out);
int ldpc_b200_decoder_set_engine(ldpc_b200_decoder_t *dec, int engine);
int ldpc_b200_decoder_engine(const ldpc_b200_decoder_t *dec);   /* engine a decode call will use */
void ldpc_b200_decoder_destroy(ldpc_b200_decoder_t *dec);

/* Host-buffer decode. llr: n_frames x N float64 (zeros at punctured positions); bits: n_frames x N, 0/1;
 * iters / converged (n_frames each) may be NULL. Synchronous. n_frames may exceed max_batch. */
int ldpc_b200_decode(ldpc_b200_decoder_t *dec, int max_iters, int n_frames, const double *llr,
                     uint8_t *bits, int32_t *iters, uint8_t *converged);
/* Same with float32 LLRs in and bit-packed decisions out (bit j of frame f = word f*ceil(N/32)+j/32, bit j%32). */
int ldpc_b200_decode_f32(ldpc_b200_decoder_t *dec, int max_iters, int n_frames, const float *llr,
                         uint32_t *bits_packed, int32_t *iters, uint8_t *converged);
/* The decode closure of mkLDPC in one pass (src/ECC/Code/LDPC/Utils.hs:55,68-72): rx holds, per frame, `take`
 * received LLRs at a row stride of rx_stride doubles; the N - take punctured positions get LLR 0 on the device and never
 * cross PCIe; msg_bits receives the first k_take decisions of every frame (n_frames x k_take bytes). */
int ldpc_b200_decode_punctured(ldpc_b200_decoder_t *dec, int max_iters, int n_frames, const double *rx, int rx_stride,
                               int take, uint8_t *msg_bits, int k_take);

/* One frame, with the messages of every executed iteration t (0-based): c2v[t*E + e] = ne'[edge e] in
 * reference edge order (row-major, ascending column), lam[t*N + j] = lam'[j]. Arrays hold max_iters rows.
 * Runs the engine the decoder decodes with (ldpc_b200_decoder_engine): the STREAM kernels, the message-tap instance of the
 * compiled RESIDENT kernel (registers dumped in reference edge order) or the ONCHIP kernel with its tap enabled; values of
 * a float32 decoder are the float32 messages widened to double. */
int ldpc_b200_decode_trace(ldpc_b200_decoder_t *dec, int max_iters, const double *llr, double *c2v,
                           double *lam, int32_t *iters, uint8_t *bits, uint8_t *converged);

/* Device-resident decode: d_llr is n_frames x N in the decoder's precision (float for FP32, double for
 * FP64), d_bits_packed is n_frames x ceil(N/32) uint32, d_iters / d_converged may be NULL.
 * early_exit = 0 executes exactly max_iters iterations for every frame (the convergence test still
 * selects the output, so decisions equal the early-exit run's). n_frames <= max_batch. Asynchronous
 * on the decoder's stream; ldpc_b200_decoder_sync waits for it. */
int ldpc_b200_decode_device(ldpc_b200_decoder_t *dec, int max_iters, int early_exit, int n_frames,
                            const void *d_llr, uint32_t *d_bits_packed, int32_t *d_iters, uint8_t *d_converged);
int ldpc_b200_decoder_sync(ldpc_b200_decoder_t *dec);
/* Host-buffer calls of one to four frames on a RESIDENT / ONCHIP decoder -- the `decode` of the reference's tester, one
 * frame per call from several threads (Utils.hs:62-69) -- go through a MAILBOX shared by all decoder handles of the code:
 * slots in mapped host memory, frames of concurrent callers coalesced into one launch of a latency-oriented kernel, no
 * copies and no stream synchronisation.  Same results as the batched path.  LDPC_B200_MAILBOX=0 in the environment
 * disables it.  This reports how many launches served how many frames through this decoder's mailbox so far. */
int ldpc_b200_decoder_mailbox_stats(const ldpc_b200_decoder_t *dec, uint64_t *launches, uint64_t *frames);
/* Device time (ms) of the decode kernels of the last decode* call on this decoder, and launches it made. */
int ldpc_b200_decoder_last_timing(ldpc_b200_decoder_t *dec, double *kernel_ms, int64_t *launches);

/* Channel frames on the device: BPSK (bit b -> 2b-1) + AWGN, sigma^2 = 1/(2 R 10^(ebn0/10)), R = k_info/n_tx,
 * LLR = 2y/sigma^2 for the first n_tx positions and 0 for the punctured rest (Utils.hs:55). Noise is
 * Philox4x32-10 keyed by seed with the global frame index as the stream id, so results do not depend on
 * batching or on the number of GPUs. d_codewords: NULL = all-zero codeword, else n_frames x ceil(N/32) packed.
 * d_llr: n_frames x N in the decoder's precision. */
int ldpc_b200_generate_llr(ldpc_b200_decoder_t *dec, double ebn0_db, int k_info, int n_tx, uint64_t seed,
                           uint64_t first_frame, int n_frames, const uint32_t *d_codewords, void *d_llr);

/* Generate + decode + count on the device for frames [first_frame, first_frame + n_frames).
 * Counters are ADDED to *out. all-zero codeword. */
int ldpc_b200_simulate(ldpc_b200_decoder_t *dec, int max_iters, int early_exit, double ebn0_db, int k_info,
                       int n_tx, uint64_t seed, uint64_t first_frame, uint64_t n_frames, ldpc_b200_counters_t *out);
/* Same, but the five integer counters are accumulated into d_counters (5 x uint64 in device memory, e.g. the
 * send buffer of an all-reduce) on the decoder's stream, with no host synchronisation. */
int ldpc_b200_simulate_device(ldpc_b200_decoder_t *dec, int max_iters, int early_exit, double ebn0_db, int k_info,
                              int n_tx, uint64_t seed, uint64_t first_frame, uint64_t n_frames, uint64_t *d_counters);

/* Single-process multi-GPU form (one host process driving n_gpus devices 0..n_gpus-1, e.g. the Haskell tester): H is
 * replicated, GPU g simulates a contiguous share of [first_frame, first_frame+n_frames), and the five counters are summed
 * with one NCCL all-reduce (libnccl.so.2 is loaded at run time). Results are independent of n_gpus. Counters are ADDED to
 * *out; kernel_ms is the maximum over the GPUs. The replicas are cached until ldpc_b200_simulate_multi_release(). */
int ldpc_b200_simulate_multi(const ldpc_b200_code_t *code, int n_gpus, int algo, int precision, int max_batch, int max_iters,
                             int early_exit, double ebn0_db, int k_info, int n_tx, uint64_t seed, uint64_t first_frame,
                             uint64_t n_frames, ldpc_b200_counters_t *out);
void ldpc_b200_simulate_multi_release(void);

/* Error counting for device-resident decisions: adds {frames, frame errors, bit errors over the first k_info bits,
 * sum of iterations, unconverged} of n_frames frames to d_counters (5 x uint64, device). d_ref_packed = transmitted
 * codewords (n_frames x ceil(N/32)) or NULL for the all-zero codeword. Asynchronous on the decoder's stream. */
int ldpc_b200_count_errors(ldpc_b200_decoder_t *dec, int n_frames, int k_info, const uint32_t *d_bits_packed,
                           const uint32_t *d_ref_packed, const int32_t *d_iters, const uint8_t *d_converged,
                           uint64_t *d_counters);

/* Raw device memory helpers so a host without a CUDA binding (Haskell, ctypes) can stage buffers. */
int ldpc_b200_malloc(void **d_ptr, uint64_t bytes);
int ldpc_b200_free(void *d_ptr);
int ldpc_b200_malloc_host(void **h_ptr, uint64_t bytes);   /* pinned */
int ldpc_b200_free_host(void *h_ptr);
int ldpc_b200_memcpy_h2d(void *d_dst, const void *h_src, uint64_t bytes);
int ldpc_b200_memcpy_d2h(void *h_dst, const void *d_src, uint64_t bytes);
int ldpc_b200_memset(void *d_ptr, int value, uint64_t bytes);
/* CUDA stream handle (cudaStream_t) of a decoder, for callers that do share a CUDA runtime. */
void *ldpc_b200_decoder_stream(ldpc_b200_decoder_t *dec);

/* Philox4x32-10 block function, exported for known-answer tests. */
int ldpc_b200_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
