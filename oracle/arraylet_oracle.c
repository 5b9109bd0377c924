/*
 * arraylet_oracle.c -- THIRD CPU statement of the reference decoders: the quasi-cyclic "arraylet" modules
 * `ldpc/arraylet-min` (src/ECC/Code/LDPC/Fast/ArrayletMin.hs) and `ldpc/arraylet` (Fast/Arraylet.hs), transcribed
 * in their own data structure -- a matrix of circulants, each holding one vector of `sz` messages -- and NOT in the
 * edge-list form of ldpc_oracle.c.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT (same rules as ldpc_oracle.c: only tests/, smoke() and bench.py's CPU
 * legs may load it).  PARITY UNPINNED by reference tests (the reference has none); this file exists to tighten the
 * pin: it is written from different reference modules than ldpc_oracle.c / dense_ref.py (which follow
 * Reference/Orig.hs and Reference/Min.hs), shares no code with them, and must agree with them on decisions and
 * iteration counts (the posteriors differ by the association of the variable update only, see `lam'` below).
 *
 * Paths relative to /root/reference/src/ECC/Code/LDPC.  Arithmetic: IEEE float64, build with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* Fast/ArrayletMin.hs:23  data Arraylet a = Arraylet Int (U.Vector a)
 * "a small array that acts like a NxN matrix with one element in each row and column"; the Int is the rotation:
 * element r of the vector sits at (row r, column (r + off) mod sz)   (lookupArraylet :33-40, arrayArraylet :42-43). */
typedef struct { int present; int off; double *v; } arraylet;

/* Fast/ArrayletMin.hs:63  data Matrixlet a = Matrixlet Int (Matrix (Maybe (Arraylet a))) */
typedef struct {
    int sz, brows, bcols;
    arraylet *a;             /* [brows * bcols], row-major; present = 0 is Nothing */
} matrixlet;

typedef struct arraylet_code {
    int sz, brows, bcols, M, N, E;
    int *off;                /* [brows * bcols] rotation, -1 = zero block */
    int *edge_base;          /* [brows * bcols] reference edge index of element 0 of the circulant, see edge_of() */
    int *row_deg;            /* [brows] circulants per block row */
} arraylet_code;

/* initMatrixlet (ArrayletMin.hs:66-77): value 2^k -> Arraylet k (replicate sz zero); 0 -> Nothing.  `shifts` is the
 * bit index k per block (-1 for a zero block), exactly what `g` computes at :75-77. */
arraylet_code *arraylet_code_create(int sz, int brows, int bcols, const int32_t *shifts)
{
    arraylet_code *c = (arraylet_code *)calloc(1, sizeof *c);
    c->sz = sz; c->brows = brows; c->bcols = bcols; c->M = brows * sz; c->N = bcols * sz;
    c->off = (int *)malloc(sizeof(int) * brows * bcols);
    c->edge_base = (int *)malloc(sizeof(int) * brows * bcols);
    c->row_deg = (int *)calloc(brows, sizeof(int));
    int E = 0;
    for (int r = 0; r < brows; r++) {
        for (int b = 0; b < bcols; b++) { c->off[r * bcols + b] = shifts[r * bcols + b]; if (shifts[r * bcols + b] >= 0) c->row_deg[r]++; }
        int k = 0;
        for (int b = 0; b < bcols; b++) {
            /* reference edge order = row-major, ascending column: check row (r, i) owns edges
             * E_before + i * row_deg[r] + (rank of the block column among the non-zero blocks of block row r) */
            c->edge_base[r * bcols + b] = shifts[r * bcols + b] >= 0 ? E + k++ : -1;
        }
        E += c->row_deg[r] * sz;
    }
    c->E = E;
    return c;
}

void arraylet_code_destroy(arraylet_code *c)
{
    if (!c) return;
    free(c->off); free(c->edge_base); free(c->row_deg); free(c);
}

int arraylet_code_edges(const arraylet_code *c) { return c->E; }

static inline int edge_of(const arraylet_code *c, int r, int b, int i) { return c->edge_base[r * c->bcols + b] + i * c->row_deg[r]; }

static matrixlet mlet_new(const arraylet_code *c)
{
    matrixlet m; m.sz = c->sz; m.brows = c->brows; m.bcols = c->bcols;
    m.a = (arraylet *)calloc((size_t)c->brows * c->bcols, sizeof(arraylet));
    for (int q = 0; q < c->brows * c->bcols; q++)
        if (c->off[q] >= 0) { m.a[q].present = 1; m.a[q].off = c->off[q]; m.a[q].v = (double *)calloc(c->sz, sizeof(double)); }
    return m;
}

static void mlet_free(matrixlet *m)
{
    for (int q = 0; q < m->brows * m->bcols; q++) free(m->a[q].v);
    free(m->a);
}

static inline double hs_signum(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }
static inline double hs_min(double x, double y) { return x <= y ? x : y; }
/* ArrayletMin.hs:188-189  min' x y = signum x * signum y * min (abs x) (abs y) */
static inline double min_prime(double x, double y) { return hs_signum(x) * hs_signum(y) * hs_min(fabs(x), fabs(y)); }
/* Utils.hs:113-117 */
static inline double atanh_prime(double x)
{
    double y = atanh(x);
    return isinf(y) ? hs_signum(x) * 18.714973875118524 : y;
}
static inline int hard(double x) { return x > 0; }

/* ans = foldColsMatrixlet (/=) $ matrixMatrixlet mLet $ \ (r,c) -> hard (lam U.! c)      (ArrayletMin.hs:179-180)
 *   matrixMatrixlet (:79-85): circulant (R, C) with rotation off becomes the vector [ k (R*sz + r', C*sz + (r'+off) mod sz) | r' ]
 *   foldColsMatrixlet (:92-96): per block row, foldr1 (zipWith f) over the present circulants in ascending block column
 * U.all (== False) ans  <=>  every check is satisfied. */
static int syndrome_ok(const arraylet_code *c, const double *lam, uint8_t *acc, uint8_t *cur)
{
    const int sz = c->sz;
    for (int R = 0; R < c->brows; R++) {
        int have = 0;
        for (int C = c->bcols - 1; C >= 0; C--) {                 /* foldr1: start from the last present circulant */
            const int off = c->off[R * c->bcols + C];
            if (off < 0) continue;
            for (int r = 0; r < sz; r++) cur[r] = (uint8_t)hard(lam[C * sz + (r + off) % sz]);
            if (!have) { memcpy(acc, cur, sz); have = 1; }
            else for (int r = 0; r < sz; r++) acc[r] = (uint8_t)(cur[r] != acc[r]);       /* zipWith (/=) x acc */
        }
        if (have) for (int r = 0; r < sz; r++) if (acc[r]) return 0;
    }
    return 1;
}

/* One iteration, ArrayletMin.hs:182-192 (algo 1) / Arraylet.hs:176-186 (algo 0):
 *   ne_tanh = foldColsMatrixlet' (\ (m,n) v -> [(n, - (lam!n - v))]) (++) ne            -- per check row, ascending column
 *             (Arraylet.hs:176-177:               [(n, tanh (- ((lam!n - v) / 2)))])
 *   ne'     = matrixMatrixlet ne $ \ (m,n) -> (-3/4) * foldr1 min' [ v | (j,v) <- ne_tanh ! m, j /= n ]
 *             (Arraylet.hs:179-180:               -2 * atanh' (product [ v | (j,v) <- ne_tanh ! m, j /= n ]))
 *   lam'    = U.zipWith (+) orig_lam $ foldRowsMatrixlet (+) ne'                          */
static void iterate(const arraylet_code *c, int algo, const double *orig_lam, const double *lam, const matrixlet *ne,
                    matrixlet *ne2, double *lam2, int *lst_n, double *lst_v, double *colacc, double *rot)
{
    const int sz = c->sz, bc = c->bcols;
    for (int R = 0; R < c->brows; R++) {
        for (int r = 0; r < sz; r++) {
            /* ne_tanh ! (R*sz + r): the list built by (++) over ascending block columns (foldColsArraylet' :55-57) */
            int d = 0;
            for (int C = 0; C < bc; C++) {
                const arraylet *x = &ne->a[R * bc + C];
                if (!x->present) continue;
                const int n = C * sz + (r + x->off) % sz;
                const double v = x->v[r];
                lst_n[d] = n;
                lst_v[d] = algo == 0 ? tanh(-((lam[n] - v) / 2)) : -((lam[n] - v));
                d++;
            }
            for (int C = 0; C < bc; C++) {
                const arraylet *x = &ne->a[R * bc + C];
                if (!x->present) continue;
                const int n = C * sz + (r + x->off) % sz;
                double out;
                if (algo == 0) {
                    double p = 1.0;                                          /* product = foldl (*) 1 */
                    for (int q = 0; q < d; q++) if (lst_n[q] != n) p = p * lst_v[q];
                    out = -2 * atanh_prime(p);
                } else {
                    double acc = 0; int have = 0;                            /* foldr1 min' over the kept elements */
                    for (int q = d - 1; q >= 0; q--) if (lst_n[q] != n) {
                        if (!have) { acc = lst_v[q]; have = 1; } else acc = min_prime(lst_v[q], acc);
                    }
                    out = (-3.0 / 4.0) * acc;
                }
                ne2->a[R * bc + C].v[r] = out;
            }
        }
    }
    /* foldRowsMatrixlet (+) ne'  (ArrayletMin.hs:104-108): per block column, foldr1 (zipWith (+)) over the present
     * circulants in ascending block row of foldRowsArraylet x (:43-45): the vector rotated so that it is indexed by
     * COLUMN, (a, b) = splitAt (len - off) m ; b ++ a. */
    for (int C = 0; C < bc; C++) {
        int have = 0;
        for (int R = c->brows - 1; R >= 0; R--) {
            const arraylet *x = &ne2->a[R * bc + C];
            if (!x->present) continue;
            const int cut = sz - x->off;                                     /* a = v[0 .. cut), b = v[cut .. sz) */
            memcpy(rot, x->v + cut, sizeof(double) * (sz - cut));
            memcpy(rot + (sz - cut), x->v, sizeof(double) * cut);
            if (!have) { memcpy(colacc, rot, sizeof(double) * sz); have = 1; }
            else for (int j = 0; j < sz; j++) colacc[j] = rot[j] + colacc[j];              /* zipWith (+) x acc */
        }
        /* "Assumes there is always one value on every column" (:91); a column without a circulant would make foldr1 fail */
        for (int j = 0; j < sz; j++) lam2[C * sz + j] = orig_lam[C * sz + j] + (have ? colacc[j] : 0.0);
    }
}

/* ldpc (ArrayletMin.hs:167-177 = Arraylet.hs:164-174):
 *   loop n ne lam | U.all (== False) ans = lam | n >= maxIterations = orig_lam | otherwise = loop (n+1) ne' lam'
 * algo: 0 = `ldpc/arraylet` (sum-product), 1 = `ldpc/arraylet-min` (alpha = 3/4 min-sum).
 * trace_c2v (max_iters x E, REFERENCE EDGE ORDER: row-major, ascending column) / trace_lam (max_iters x N) may be NULL. */
int arraylet_decode(const arraylet_code *c, int algo, int max_iters, const double *llr, uint8_t *bits, int *converged,
                    double *trace_c2v, double *trace_lam)
{
    const int sz = c->sz, N = c->N;
    matrixlet ne = mlet_new(c), ne2 = mlet_new(c);                  /* mLet = initMatrixlet 0 a */
    double *lam = (double *)malloc(sizeof(double) * N), *lam2 = (double *)malloc(sizeof(double) * N);
    int *lst_n = (int *)malloc(sizeof(int) * c->bcols);
    double *lst_v = (double *)malloc(sizeof(double) * c->bcols);
    double *colacc = (double *)malloc(sizeof(double) * sz), *rot = (double *)malloc(sizeof(double) * sz);
    uint8_t *sa = (uint8_t *)malloc(sz), *sb = (uint8_t *)malloc(sz);
    memcpy(lam, llr, sizeof(double) * N);
    int n = 0, ok;
    for (;;) {
        ok = syndrome_ok(c, lam, sa, sb);
        if (ok || n >= max_iters) break;
        iterate(c, algo, llr, lam, &ne, &ne2, lam2, lst_n, lst_v, colacc, rot);
        if (trace_c2v)
            for (int R = 0; R < c->brows; R++)
                for (int C = 0; C < c->bcols; C++)
                    if (ne2.a[R * c->bcols + C].present)
                        for (int r = 0; r < sz; r++) trace_c2v[(size_t)n * c->E + edge_of(c, R, C, r)] = ne2.a[R * c->bcols + C].v[r];
        if (trace_lam) memcpy(trace_lam + (size_t)n * N, lam2, sizeof(double) * N);
        matrixlet t = ne; ne = ne2; ne2 = t;
        double *tl = lam; lam = lam2; lam2 = tl;
        n++;
    }
    const double *res = ok ? lam : llr;                            /* U.map hard $ loop ... */
    for (int j = 0; j < N; j++) bits[j] = (uint8_t)hard(res[j]);
    if (converged) *converged = ok;
    mlet_free(&ne); mlet_free(&ne2);
    free(lam); free(lam2); free(lst_n); free(lst_v); free(colacc); free(rot); free(sa); free(sb);
    return n;
}

typedef struct {
    const arraylet_code *c; int algo, max_iters, n_frames;
    const double *llr; uint8_t *bits; int32_t *iters; uint8_t *converged;
    int next; pthread_mutex_t mu;
} al_job;

static void *al_worker(void *arg)
{
    al_job *j = (al_job *)arg;
    const int N = j->c->N;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int f = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (f >= j->n_frames) break;
        int cv = 0;
        int it = arraylet_decode(j->c, j->algo, j->max_iters, j->llr + (size_t)f * N, j->bits + (size_t)f * N, &cv, NULL, NULL);
        if (j->iters) j->iters[f] = it;
        if (j->converged) j->converged[f] = (uint8_t)cv;
    }
    return NULL;
}

int arraylet_decode_batch(const arraylet_code *c, int algo, int max_iters, int n_frames, const double *llr, uint8_t *bits,
                          int32_t *iters, uint8_t *converged, int n_threads)
{
    if (n_threads <= 0) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    al_job j = { c, algo, max_iters, n_frames, llr, bits, iters, converged, 0, PTHREAD_MUTEX_INITIALIZER };
    pthread_t th[256];
    for (int t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, al_worker, &j);
    al_worker(&j);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
    return n_threads;
}
