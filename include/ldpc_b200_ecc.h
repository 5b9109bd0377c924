/*
 * ldpc_b200_ecc.h -- host layer of libldpc_b200.so above the decoder ABI (ldpc_b200.h): what the reference builds
 * around its decoders so that a tester can drive them.  Plain C ABI, written in C++ (the reference's host is
 * compiled Haskell; no GHC exists in this image).
 *
 * Reference interfaces mirrored (paths relative to /root/reference):
 *
 *   ldpc_b200_matrix_load     loadMatrix: search order q / alist / m under <root>/ (src/Data/BitMatrix/Loader.hs:58-81),
 *                             .q semantics (src/Data/Matrix/QuasiCyclic.hs:19-25,52-56), .alist in the reference's
 *                             orientation with zeros dropped (src/Data/BitMatrix/Alist.hs:30-46), .m 0/1 rows
 *                             (src/Data/BitMatrix/Matlab.hs:19-26)
 *   ldpc_b200_ecc_open        the matcher of mkLDPC_CodeIO + mkLDPC (src/ECC/Code/LDPC/Utils.hs:35-75,100-108):
 *                             "ldpc/<decoder>/<matrix-name>/<max-rounds>[/<num>/<den>]" -> ECC record
 *   ldpc_b200_ecc_name / _message_length / _codeword_length     fields of the record (Utils.hs:59-75)
 *   ldpc_b200_ecc_encode      `encode`: inp ++ take (c_length - m_length) (encoder' inp)  (Utils.hs:61), with the
 *                             circulant-word multiply of src/ECC/Code/LDPC/Fast/Encoder.hs:26-63 for .q generators and
 *                             the dense v*G of src/ECC/Code/LDPC/Reference/Orig.hs:25-26 otherwise
 *   ldpc_b200_ecc_decode      `decode`: unpuncture, decoder, take m_length, flag (Utils.hs:55,62-72)
 *   ldpc_b200_ecc_run         one Eb/N0 point of ecc-manifold's tester (external): frames on the
 *                             device (random messages through the device encoder, or the all-zero codeword), BPSK +
 *                             AWGN, decode, count errors over the first message_length bits
 *   ldpc_b200_ber_line        the result line of eccPrinter as visible in NOTES.txt:3
 *
 * Decoder names registered: b200-minsum, b200-spa (fp32), b200-minsum-fp64, b200-spa-fp64, b200-spa-cm-fp64
 * (the cached-product variant of src/ECC/Code/LDPC/Fast/CachedMult.hs).
 */
#ifndef LDPC_B200_ECC_H
#define LDPC_B200_ECC_H

#include <stdint.h>
#include "ldpc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_B200_ENOMATCH (-6)   /* the name is not one of ours: the reference's matcher returns [] */
#define LDPC_B200_ENOENT (-7)     /* "can not find any matrix files in ..." (Loader.hs:73) */

typedef struct ldpc_b200_matrix ldpc_b200_matrix_t;   /* a loaded bit matrix: CSR rows + optional circulant words */
typedef struct ldpc_b200_ecc ldpc_b200_ecc_t;

/* ---- matrix files ---- */
/* path_noext e.g. "jpl.4096.4.5/H"; tries <root>/<path>.q, .alist, .m in that order. */
int ldpc_b200_matrix_load(const char *codes_root, const char *path_noext, ldpc_b200_matrix_t **out);
/* Parse from memory; kind = "q", "alist" or "m". */
int ldpc_b200_matrix_parse(const char *kind, const char *text, ldpc_b200_matrix_t **out);
void ldpc_b200_matrix_destroy(ldpc_b200_matrix_t *m);
/* rows / cols of the EXPANDED matrix, number of ones, circulant size (0 if not quasi-cyclic). */
int ldpc_b200_matrix_info(const ldpc_b200_matrix_t *m, int *rows, int *cols, int64_t *ones, int *qc_sz);
/* CSR copy-out: row_ptr has rows+1 entries, col_idx `ones` entries (ascending within a row). */
int ldpc_b200_matrix_csr(const ldpc_b200_matrix_t *m, int32_t *row_ptr, int32_t *col_idx);
/* For single-weight quasi-cyclic matrices: shifts[brows*bcols] (-1 = zero block); fails otherwise. */
int ldpc_b200_matrix_qc_shifts(const ldpc_b200_matrix_t *m, int *brows, int *bcols, int32_t *shifts);

/* ---- ECC record ---- */
/* Returns LDPC_B200_ENOMATCH when `name` is not in our grammar. max_batch sizes the device buffers of the decoder
 * (1 is enough for the one-frame-per-call use of the reference's tester). No GPU is touched until the first decode. */
int ldpc_b200_ecc_open(const char *codes_root, const char *name, int max_batch, ldpc_b200_ecc_t **out);
void ldpc_b200_ecc_close(ldpc_b200_ecc_t *e);
const char *ldpc_b200_ecc_name(const ldpc_b200_ecc_t *e);
int ldpc_b200_ecc_message_length(const ldpc_b200_ecc_t *e);
int ldpc_b200_ecc_codeword_length(const ldpc_b200_ecc_t *e);
int ldpc_b200_ecc_unpunctured_length(const ldpc_b200_ecc_t *e);
int ldpc_b200_ecc_max_iters(const ldpc_b200_ecc_t *e);
/* msg: message_length bits (0/1 bytes); cw: codeword_length bits. Host side, no GPU. */
int ldpc_b200_ecc_encode(const ldpc_b200_ecc_t *e, const uint8_t *msg, uint8_t *cw);
/* rx: n_frames x rx_len received LLRs (the tester hands over codeword_length of them); out: n_frames x message_length
 * bits; flags (n_frames, may be NULL) = 1: every decoder of the reference returns Just. */
int ldpc_b200_ecc_decode(ldpc_b200_ecc_t *e, int n_frames, const double *rx, int rx_len, uint8_t *out, uint8_t *flags);
/* One Eb/N0 point: n_frames frames [first_frame, first_frame+n_frames) generated, decoded and counted on the device.
 * random_messages = 1 draws a message per frame (Philox) and encodes it on the device; 0 sends the all-zero codeword.
 * Counters are ADDED to *out. */
int ldpc_b200_ecc_run(ldpc_b200_ecc_t *e, double ebn0_db, uint64_t seed, uint64_t first_frame, uint64_t n_frames,
                      int random_messages, ldpc_b200_counters_t *out);
/* "<secs> <name> <EbN0> <#frames> <#bit-errors> <BER> +<p>% -<m>%": the tester's line (NOTES.txt:3), with a 95 %
 * Wilson interval on the bit error rate. Returns the number of characters written (snprintf semantics). */
int ldpc_b200_ber_line(char *buf, int buf_len, double seconds, const char *name, double ebn0_db,
                       uint64_t frames, uint64_t bit_errors, int message_length);

/* The reference tester's calling pattern as a measurable unit (Utils.hs:53,62-69): `n_threads` tester threads,
 * thread t owning decoder t (one ECC record each, opened like mkLDPC's `replicateM maxThreadCount (decoder0 h)`), every thread
 * calling ECC.decode with ONE received vector per call for its share of the frames (frame f belongs to thread f mod
 * n_threads).  rx: n_frames x rx_len; out: n_frames x message_length.  *seconds = wall time from the common start of the
 * threads to the last join (records are opened and warmed before the clock starts).  mailbox_launches / mailbox_frames
 * (may be NULL): how many kernel launches served the calls. */
int ldpc_b200_ecc_tester_pattern(const char *codes_root, const char *name, int n_threads, int n_frames, const double *rx, int rx_len,
                                 uint8_t *out, double *seconds, uint64_t *mailbox_launches, uint64_t *mailbox_frames);

/* Device encoder on packed words: d_msg n_frames x ceil(K/32) -> d_cw n_frames x ceil(N/32) with the
 * parity bits beyond codeword_length cleared. Asynchronous on the ECC's decoder stream. */
int ldpc_b200_ecc_encode_device(ldpc_b200_ecc_t *e, int n_frames, const uint32_t *d_msg_packed, uint32_t *d_cw_packed);
/* Wait for the asynchronous work of this ECC's decoder stream. */
int ldpc_b200_ecc_sync(ldpc_b200_ecc_t *e);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_ECC_H */
