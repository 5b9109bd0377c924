/*
 * ldpc_oracle.c -- CPU restatement of ku-fpg/ecc-ldpc's belief-propagation decoders.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the product path
 * (libldpc_b200.so) never links, loads or calls anything in oracle/.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer fixtures for
 * this path and cannot be compiled here (no GHC, ecc-manifold absent).
 * The pins we do have are constructed: a second, independent dense numpy statement
 * (oracle/dense_ref.py), the fixtures it generated (tests/golden/), and domain identities.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Arithmetic is IEEE float64 with the reference's operation order;
 * build with -ffp-contract=off so no FMA contraction changes the rounding.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef struct oracle_code {
    int M, N, E;
    int *row_ptr;   /* M+1 : edges of check m are [row_ptr[m], row_ptr[m+1]), ascending column */
    int *col_idx;   /* E   : variable index of each edge (reference edge order: row-major, ascending col) */
    int *col_ptr;   /* N+1 */
    int *col_edge;  /* E   : edge ids of variable j in ascending row order */
} oracle_code;

/* H as a sparse matrix; the row-major / ascending-column edge order is the order in which the
 * dense generators of Reference/Orig.hs:82-92 visit the ones of H. */
oracle_code *oracle_code_create(int M, int N, const int32_t *row_ptr, const int32_t *col_idx)
{
    oracle_code *c = (oracle_code *)calloc(1, sizeof *c);
    int E = row_ptr[M];
    c->M = M; c->N = N; c->E = E;
    c->row_ptr = (int *)malloc(sizeof(int) * (M + 1));
    c->col_idx = (int *)malloc(sizeof(int) * (E > 0 ? E : 1));
    c->col_ptr = (int *)calloc(N + 1, sizeof(int));
    c->col_edge = (int *)malloc(sizeof(int) * (E > 0 ? E : 1));
    memcpy(c->row_ptr, row_ptr, sizeof(int) * (M + 1));
    memcpy(c->col_idx, col_idx, sizeof(int) * E);
    for (int e = 0; e < E; e++) c->col_ptr[col_idx[e] + 1]++;
    for (int j = 0; j < N; j++) c->col_ptr[j + 1] += c->col_ptr[j];
    int *fill = (int *)calloc(N, sizeof(int));
    for (int m = 0; m < M; m++)           /* ascending m => ascending row order inside each column */
        for (int e = row_ptr[m]; e < row_ptr[m + 1]; e++) {
            int j = col_idx[e];
            c->col_edge[c->col_ptr[j] + fill[j]++] = e;
        }
    free(fill);
    return c;
}

void oracle_code_destroy(oracle_code *c)
{
    if (!c) return;
    free(c->row_ptr); free(c->col_idx); free(c->col_ptr); free(c->col_edge); free(c);
}

int oracle_code_edges(const oracle_code *c) { return c->E; }

/* Haskell's signum on Double: 1 / -1 / x itself for +-0 and NaN (GHC.Float). */
static inline double hs_signum(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }
/* Haskell's min on Double: if x <= y then x else y. */
static inline double hs_min(double x, double y) { return x <= y ? x : y; }

/* Reference/Min.hs:89  min' x y = signum x * signum y * min (abs x) (abs y) */
static inline double min_prime(double x, double y)
{
    return hs_signum(x) * hs_signum(y) * hs_min(fabs(x), fabs(y));
}

/* Utils.hs:113-117  atanh' : atanh, but +-inf is replaced by +-18.714973875118524 */
static inline double atanh_prime(double x)
{
    double y = atanh(x);
    if (isinf(y)) return hs_signum(x) * 18.714973875118524;
    return y;
}

/* hard decision: ecc-manifold's `hard`, restated in GPU/Reference.hs:59-60 as (> 0). */
static inline int hard(double x) { return x > 0; }

/* Orig.hs:73-78 / Min.hs:69-73: c_hat = hard lam ; ans = H * c_hat over GF(2); all False? */
static int syndrome_ok(const oracle_code *c, const double *lam)
{
    for (int m = 0; m < c->M; m++) {
        int p = 0;
        for (int e = c->row_ptr[m]; e < c->row_ptr[m + 1]; e++) p ^= hard(lam[c->col_idx[e]]);
        if (p) return 0;
    }
    return 1;
}

/* One check-node update in the literal form of the reference (O(d^2) per row).
 * SPA: Orig.hs:81-92   ne'[m,n] = -2 * atanh'(product [ tanh(-((lam_j - ne[m,j])/2)) | j asc, j /= n ])
 * MS : Min.hs:75-87    ne'[m,n] = (-3/4) * foldr1 min' [ -(lam_j - ne[m,j]) | j asc, j /= n ]       */
static void cn_update_literal(const oracle_code *c, int algo, const double *lam, const double *ne, double *ne2)
{
    double t[4096];
    for (int m = 0; m < c->M; m++) {
        int b = c->row_ptr[m], d = c->row_ptr[m + 1] - b;
        for (int k = 0; k < d; k++) {
            double v = lam[c->col_idx[b + k]] - ne[b + k];
            t[k] = (algo == 0) ? tanh(-(v / 2)) : -v;
        }
        for (int k = 0; k < d; k++) {
            if (algo == 0) {
                double p = 1.0;                       /* product = foldl (*) 1 */
                for (int q = 0; q < d; q++) if (q != k) p = p * t[q];
                ne2[b + k] = -2 * atanh_prime(p);
            } else {
                /* foldr1 f [a0..an] = f a0 (f a1 (... an)) : start from the last kept element */
                double acc = 0; int have = 0;
                for (int q = d - 1; q >= 0; q--) if (q != k) {
                    if (!have) { acc = t[q]; have = 1; } else acc = min_prime(t[q], acc);
                }
                ne2[b + k] = (-3.0 / 4.0) * acc;      /* d == 1 would be foldr1 [] = error in Haskell */
            }
        }
    }
}

/* Fast/CachedMult.hs:25-56: StableDiv = (the value closest to zero, the product of all the others).
 *   lit x | x >= 1 = (1, x) | otherwise = (x, 1)
 *   smult (a,b) (c,d) = (minOfMins, b * maxOfMins * d)  where (minOfMins, maxOfMins) = absMinMax a c
 *   absMinMax x y | abs x < abs y = (x, y) | otherwise = (y, x)
 *   sdiv (a,b) c | a == c = b | otherwise = a * (b / c)                                         */
typedef struct { double a, b; } stable_div;
static inline stable_div sd_lit(double x) { stable_div r; if (x >= 1) { r.a = 1; r.b = x; } else { r.a = x; r.b = 1; } return r; }
static inline stable_div sd_mult(stable_div p, stable_div q)
{
    double mn, mx;
    if (fabs(p.a) < fabs(q.a)) { mn = p.a; mx = q.a; } else { mn = q.a; mx = p.a; }
    stable_div r; r.a = mn; r.b = (p.b * mx) * q.b;
    return r;
}
static inline double sd_div(stable_div p, double c) { return p.a == c ? p.b : p.a * (p.b / c); }

/* Fast/CachedMult.hs:248-261 (`ldpc/arraylet-cm`): sum-product whose leave-one-out product is
 * (product of the whole row) / (own term), the row product kept as a StableDiv:
 *   t[m,n]   = tanh (- ((lam_n - ne[m,n]) / 2))
 *   row m    = foldr1 smult [ lit t[m,n] | n ascending ]            (foldColsMatrixletU, :193-197)
 *   ne'[m,n] = -2 * atanh' (row m `sdiv` t[m,n])                                                  */
static void cn_update_cached_mult(const oracle_code *c, const double *lam, const double *ne, double *ne2)
{
    double t[4096];
    for (int m = 0; m < c->M; m++) {
        int b = c->row_ptr[m], d = c->row_ptr[m + 1] - b;
        if (d <= 0) continue;
        for (int k = 0; k < d; k++) t[k] = tanh(-((lam[c->col_idx[b + k]] - ne[b + k]) / 2));
        stable_div acc = sd_lit(t[d - 1]);
        for (int k = d - 2; k >= 0; k--) acc = sd_mult(sd_lit(t[k]), acc);
        for (int k = 0; k < d; k++) ne2[b + k] = -2 * atanh_prime(sd_div(acc, t[k]));
    }
}

/* Fast/CachedMult.hs:263-264 (same in Fast/Arraylet.hs:185-186, ArrayletMin.hs:191-192):
 *   lam' = zipWith (+) orig_lam (foldRowsMatrixlet (+) ne')  =  L_j + (ne'[m1] + (ne'[m2] + ... + ne'[mk])) */
static void vn_update_arraylet(const oracle_code *c, const double *L, const double *ne2, double *lam)
{
    for (int j = 0; j < c->N; j++) {
        int q0 = c->col_ptr[j], q1 = c->col_ptr[j + 1];
        if (q1 == q0) { lam[j] = L[j]; continue; }       /* (no fixture has an empty column) */
        double acc = ne2[c->col_edge[q1 - 1]];
        for (int q = q1 - 2; q >= q0; q--) acc = ne2[c->col_edge[q]] + acc;
        lam[j] = L[j] + acc;
    }
}

/* Same results as cn_update_literal for min-sum (up to the sign of an exact zero) in O(d):
 * the two-minimum + sign-parity form sketched in Model.hs:326-343 (decoder8) and Utils.hs:120-144. */
static void cn_update_minsum_twomin(const oracle_code *c, const double *lam, const double *ne, double *ne2)
{
    for (int m = 0; m < c->M; m++) {
        int b = c->row_ptr[m], d = c->row_ptr[m + 1] - b;
        double m1 = INFINITY, m2 = INFINITY; int neg = 0, zeros = 0;
        for (int k = 0; k < d; k++) {
            double x = -(lam[c->col_idx[b + k]] - ne[b + k]);
            double a = fabs(x);
            ne2[b + k] = x;
            if (a < m1) { m2 = m1; m1 = a; } else if (a < m2) m2 = a;
            neg ^= (x < 0); zeros += (x == 0);
        }
        (void)zeros;
        for (int k = 0; k < d; k++) {
            double x = ne2[b + k], a = fabs(x);
            double mag = (a == m1) ? m2 : m1;          /* ties: m2 == m1, either is right */
            int sgn = neg ^ (x < 0);
            double v = sgn ? -mag : mag;
            ne2[b + k] = (-3.0 / 4.0) * v;
        }
    }
}

/* Orig.hs:95-98 / Min.hs:101-104: lam'[j] = foldr (+) orig_lam[j] (col j of ne')
 *   = ne'[m1] + (ne'[m2] + (... + (ne'[mk] + L_j)))  for rows m1 < ... < mk; non-edges add 0.0 */
static void vn_update(const oracle_code *c, const double *L, const double *ne2, double *lam)
{
    for (int j = 0; j < c->N; j++) {
        double acc = L[j];
        for (int q = c->col_ptr[j + 1] - 1; q >= c->col_ptr[j]; q--) acc = ne2[c->col_edge[q]] + acc;
        lam[j] = acc;
    }
}

/* The driver loop: Orig.hs:58-71 == Min.hs:54-67.
 *   loop n ne lam | syndrome ok -> lam | n >= maxIterations -> orig_lam | otherwise -> next
 *   result = map hard (...)
 * algo: 0 = sum-product (Orig.hs), 1 = normalised min-sum alpha=3/4 (Min.hs),
 *       2 = sum-product with the cached row product of Fast/CachedMult.hs (same loop, CachedMult.hs:233-243),
 *       3 = `ldpc/arraylet` (Fast/Arraylet.hs:164-186) and 4 = `ldpc/arraylet-min` (Fast/ArrayletMin.hs:167-192) in
 *       edge-list form: the check updates of 0 / 1 with the arraylet association of the posterior,
 *       L_j + (ne'[m1] + (ne'[m2] + ...)); the circulant-vector transcription of the same two modules is
 *       arraylet_oracle.c and must agree with these bit for bit.
 * variant: 0 = literal O(d^2) check update, 1 = two-min O(d) (min-sum only).
 * trace_c2v (max_iters*E) / trace_lam (max_iters*N): after iteration t (0-based) the new messages /
 * posteriors are stored at offset t; may be NULL.
 * Returns iterations executed; *converged = 1 iff the syndrome test passed (else fallback to hard(L)). */
int oracle_decode(const oracle_code *c, int algo, int variant, int max_iters, const double *llr,
                  uint8_t *bits, int *converged, double *trace_c2v, double *trace_lam)
{
    int N = c->N, E = c->E;
    double *lam = (double *)malloc(sizeof(double) * N);
    double *ne = (double *)calloc(E > 0 ? E : 1, sizeof(double));      /* orig_ne = 0  (Orig.hs:64-65) */
    double *ne2 = (double *)malloc(sizeof(double) * (E > 0 ? E : 1));
    memcpy(lam, llr, sizeof(double) * N);
    int n = 0, ok;
    for (;;) {
        ok = syndrome_ok(c, lam);
        if (ok) break;
        if (n >= max_iters) break;
        const int cn_algo = algo == 3 ? 0 : (algo == 4 ? 1 : algo);
        if (algo == 2) cn_update_cached_mult(c, lam, ne, ne2);
        else if (cn_algo == 1 && variant == 1) cn_update_minsum_twomin(c, lam, ne, ne2);
        else cn_update_literal(c, cn_algo, lam, ne, ne2);
        if (algo >= 2) vn_update_arraylet(c, llr, ne2, lam);
        else vn_update(c, llr, ne2, lam);
        if (trace_c2v) memcpy(trace_c2v + (size_t)n * E, ne2, sizeof(double) * E);
        if (trace_lam) memcpy(trace_lam + (size_t)n * N, lam, sizeof(double) * N);
        double *t = ne; ne = ne2; ne2 = t;
        n++;
    }
    const double *res = ok ? lam : llr;                /* FALLBACK to the channel LLRs (Orig.hs:70) */
    for (int j = 0; j < N; j++) bits[j] = (uint8_t)hard(res[j]);
    if (converged) *converged = ok;
    free(lam); free(ne); free(ne2);
    return n;
}

/* Batch form used as the CPU baseline: frames are pulled one at a time by n_threads pthreads
 * (the reference runs one decoder per tester thread, Utils.hs:53-54,62-69). llr is n_frames x N. */
typedef struct {
    const oracle_code *c; int algo, variant, max_iters, n_frames;
    const double *llr; uint8_t *bits; int32_t *iters; uint8_t *converged;
    int next; pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    int N = j->c->N;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int f0 = j->next; j->next += 4;
        pthread_mutex_unlock(&j->mu);
        if (f0 >= j->n_frames) break;
        int f1 = f0 + 4 < j->n_frames ? f0 + 4 : j->n_frames;
        for (int f = f0; f < f1; f++) {
            int cv = 0;
            int it = oracle_decode(j->c, j->algo, j->variant, j->max_iters, j->llr + (size_t)f * N,
                                   j->bits + (size_t)f * N, &cv, NULL, NULL);
            if (j->iters) j->iters[f] = it;
            if (j->converged) j->converged[f] = (uint8_t)cv;
        }
    }
    return NULL;
}

int oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int oracle_decode_batch(const oracle_code *c, int algo, int variant, int max_iters, int n_frames,
                        const double *llr, uint8_t *bits, int32_t *iters, uint8_t *converged, int n_threads)
{
    if (n_threads <= 0) n_threads = oracle_max_threads();
    if (n_threads > 256) n_threads = 256;
    batch_job j = { c, algo, variant, max_iters, n_frames, llr, bits, iters, converged, 0, PTHREAD_MUTEX_INITIALIZER };
    pthread_t th[256];
    for (int t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, batch_worker, &j);
    batch_worker(&j);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
    return n_threads;
}

/* The mkLDPC decode closure, Utils.hs:55,62-72:
 *   unpuncture xs = take c_length xs ++ replicate (N - c_length) 0 ; run decoder ; take m_length.
 * rx has rx_len entries (the tester hands over c_length of them). out gets K bits. Returns the flag
 * (True: every decoder in the reference returns Just). */
int oracle_ecc_decode(const oracle_code *c, int algo, int variant, int max_iters, int K, int c_length,
                      const double *rx, int rx_len, uint8_t *out, int *iters, int *converged)
{
    int N = c->N;
    double *full = (double *)calloc(N, sizeof(double));
    uint8_t *bits = (uint8_t *)malloc(N);
    int take = c_length < rx_len ? c_length : rx_len;
    if (take > N) take = N;
    memcpy(full, rx, sizeof(double) * take);
    int it = oracle_decode(c, algo, variant, max_iters, full, bits, converged, NULL, NULL);
    memcpy(out, bits, K);
    if (iters) *iters = it;
    free(full); free(bits);
    return 1;
}

/* Dense encoder, Orig.hs:25-26 + Utils.hs:61: parity = v * G over GF(2) (G is K x P, row lists),
 * codeword = v ++ take (c_length - K) parity.  g_row_ptr/g_col_idx: CSR of G's parity part. */
void oracle_encode(int K, int P, const int32_t *g_row_ptr, const int32_t *g_col_idx, int c_length,
                   const uint8_t *msg, uint8_t *cw)
{
    uint8_t *par = (uint8_t *)calloc(P > 0 ? P : 1, 1);
    for (int k = 0; k < K; k++)
        if (msg[k]) for (int e = g_row_ptr[k]; e < g_row_ptr[k + 1]; e++) par[g_col_idx[e]] ^= 1;
    memcpy(cw, msg, K);
    for (int i = 0; i < c_length - K; i++) cw[K + i] = par[i];
    free(par);
}
