// stream_engine.cu -- STREAM engine: flooding BP for ANY parity-check matrix, messages in HBM.
//
// This engine states the reference algorithm in its own operation order, so the fp64 instance is the
// message-level verification mode (ldpc_b200_decode_trace); it is also the home of the cached-product variant
// (LDPC_B200_ALGO_SPA_CM) and the fall-back for codes too large for the on-chip engines.
//
// Layout (frame-minor, so every row of a block of TPB frames is one contiguous segment):
//   L  [N][Bp]   channel LLRs                 lam [N][Bp]  posteriors
//   ne [Bp/TPB][E][TPB]   check->variable messages, frame-block-major, reference edge order (row-major, ascending column)
// One thread = one frame of a (check group, frame block) CTA in the check-node kernel, of a (variable group, frame
// block) CTA in the variable-node kernel; the rows a CTA touches are staged by 1-D bulk copies (k_cn_tma / k_vn_tma;
// k_cn / k_vn are the dynamic-loop fall-backs).  Algorithmic traffic per frame and iteration is the reference's state
// model, 3E + 3N words: CN reads E ne + gathers lam, writes E ne; VN reads E ne + N L, writes N lam --
// and that is what ncu measures in DRAM.
//
// Reference lines restated: driver loop Orig.hs:58-71 / Min.hs:54-67; syndrome Orig.hs:73-78;
// SPA check update Orig.hs:81-92 with atanh' of Utils.hs:113-117; min-sum check update Min.hs:75-87;
// variable update Orig.hs:95-98 (right fold: last row first, starting from the channel LLR).
#include "common.cuh"
#include <math.h>
#include <utility>

namespace ldpc {

#include "bp_tree.cuh"
#include "tma.cuh"

#ifndef STREAM_TPB
#define STREAM_TPB 128
#endif
static constexpr int TPB = STREAM_TPB;   // threads per block along the frame axis
static constexpr int ACTIVE_SLOTS = 1024; // per-iteration "frames still running" counters (iteration n uses slot n mod ACTIVE_SLOTS)
static constexpr int SYNC_EVERY = 4;      // the host looks at the counter every SYNC_EVERY iterations (see stream_decode_t)

// The messages are stored frame-block-major, ne[frame block][E][TPB]: the d message rows of a check for one block of
// frames are then ONE contiguous chunk (one bulk copy in, one out, instead of d each way; the kernels are bound by the
// number of bulk operations, not by bytes).  L and lam stay [N][Bp]: their rows are gathered one by one anyway.
__host__ __device__ __forceinline__ size_t ne_at(int e, int f, int E)
{
    return ((size_t)(f / TPB) * (size_t)E + (size_t)e) * TPB + (size_t)(f % TPB);
}

template <typename T> struct Num;
template <> struct Num<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double tanh_(double x) { return tanh(x); }
    static __device__ __forceinline__ double atanh_(double x) { return atanh(x); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
};
template <> struct Num<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float tanh_(float x) { return tanhf(x); }
    static __device__ __forceinline__ float atanh_(float x) { return atanhf(x); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};

// llr [F][N] (frame-major, TIn) -> L, lam [N][Bp] (frame-minor, T). 32x32 tile transpose through smem.
template <typename T, typename TIn>
__global__ void k_load(const TIn *__restrict__ llr, T *__restrict__ L, T *__restrict__ lam, int N, int F, int Bp)
{
    __shared__ T tile[32][33];
    int j0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int f = f0 + r, j = j0 + threadIdx.x;
        tile[r][threadIdx.x] = (f < F && j < N) ? (T)llr[(size_t)f * N + j] : (T)0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int j = j0 + r, f = f0 + threadIdx.x;
        if (j < N && f < Bp) {
            T v = tile[threadIdx.x][r];
            L[(size_t)j * Bp + f] = v;
            lam[(size_t)j * Bp + f] = v;
        }
    }
}

// Orig.hs:73-78: c_hat = hard lam (x > 0); H * c_hat over GF(2). One thread per (check, frame).
template <typename T>
__global__ void k_syndrome(DeviceGraph g, const T *__restrict__ lam, const uint8_t *__restrict__ state,
                           uint8_t *__restrict__ unsat, int F, int Bp, int skip_done)
{
    int f = blockIdx.y * TPB + threadIdx.x, m = blockIdx.x;       // frame block = slow grid axis: see launch comment
    if (f >= F) return;
    if (skip_done && state[f]) return;
    int p = 0;
    for (int e = g.row_ptr[m]; e < g.row_ptr[m + 1]; e++) p ^= (lam[(size_t)g.col_idx[e] * Bp + f] > (T)0);
    if (p) unsat[f] = 1;
}

// Orig.hs:68-70: syndrome ok -> stop with lam; n >= maxIterations -> stop with orig_lam.
__global__ void k_status(uint8_t *__restrict__ state, uint8_t *__restrict__ unsat, int32_t *__restrict__ iters,
                         int *__restrict__ active, int n, int max_iters, int F)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    int running = 0;
    if (f < F) {
        if (state[f] == 0) {
            if (!unsat[f]) { state[f] = 1; iters[f] = n; }
            else if (n >= max_iters) { state[f] = 2; iters[f] = n; }
            else running = 1;
        }
        unsat[f] = 0;
    }
    unsigned b = __ballot_sync(0xffffffffu, running);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(active, __popc(b));
}

// Hard decision of the frames that stopped at iteration n: lam if converged, L on fallback.  It has to run in the iteration
// the frame stops: the staged kernels below move whole 128-frame rows, so a stopped frame's posteriors are overwritten by
// later launches.  CTA = (HARD_WPC words, 128 frames); most launches find no frame that stopped and leave at once, which is
// why a CTA covers several words (11 440 one-word CTAs per launch cost ~10 us of pure scheduling).
static constexpr int HARD_WPC = 8;
template <typename T>
__global__ void k_hard(const T *__restrict__ lam, const T *__restrict__ L, const uint8_t *__restrict__ state,
                       const int32_t *__restrict__ iters, uint32_t *__restrict__ bits, int n, int N, int F, int Bp)
{
    const int f = blockIdx.y * TPB + threadIdx.x;
    const uint8_t s = f < F ? state[f] : 0;
    const bool mine = s != 0 && iters[f] == n;
    if (!__syncthreads_or(mine)) return;
    if (!mine) return;
    const T *src = (s == 1) ? lam : L;
    const int W = (N + 31) / 32;
    for (int w = blockIdx.x * HARD_WPC; w < min(W, (int)(blockIdx.x + 1) * HARD_WPC); w++) {
        uint32_t word = 0;
        const int j0 = w * 32;
#pragma unroll 8
        for (int b = 0; b < 32; b++) {
            const int j = j0 + b;
            if (j < N && src[(size_t)j * Bp + f] > (T)0) word |= (1u << b);
        }
        bits[(size_t)f * W + w] = word;
    }
}

// Utils.hs:113-117
template <typename T>
__device__ __forceinline__ T atanh_prime(T x)
{
    T y = Num<T>::atanh_(x);
    if (isinf(y)) return x > (T)0 ? (T)18.714973875118524 : (T)-18.714973875118524;
    return y;
}

// Check-node update, one thread per (check, frame): fall-back form with dynamic loops (rows wider than 32, and the
// literal fp64 sum-product).  The staged kernels below are the ones that run for the codes of codes/.
template <typename T, int ALGO, int MAXD>
__global__ void k_cn(DeviceGraph g, const T *__restrict__ lam, T *__restrict__ ne,
                     const uint8_t *__restrict__ state, uint8_t *__restrict__ unsat, int F, int Bp, int skip_done)
{
    int f = blockIdx.y * TPB + threadIdx.x, m = blockIdx.x;
    if (f >= F) return;
    if (skip_done && state[f]) return;
    int b = g.row_ptr[m], d = g.row_ptr[m + 1] - b;
    {   // syndrome of the CURRENT posteriors (Orig.hs:73-78) rides on the same gather; k_status reads it after this kernel
        int par = 0;
        for (int k = 0; k < d; k++) par ^= (lam[(size_t)g.col_idx[b + k] * Bp + f] > (T)0);
        if (par) unsat[f] = 1;
    }
    if (ALGO == LDPC_B200_ALGO_SPA && sizeof(T) == 4) {
        // fp32 production form (see spa_fp32.cuh): complements c_j = 1 - tanh(|x_j|/2) combined leave-one-out by
        // prefix / suffix, magnitude 2 atanh(1 - c) capped at the reference's clamp, sign = -(product of the other signs).
        // The literal tanh-product below saturates in fp32 and shows an error floor the fp64 reference does not have.
        float c[MAXD], pre[MAXD];
        uint32_t sgn = 0;
        float run = 0.0f;                                  // identity: c = 0 <=> tanh = 1
        for (int k = 0; k < d; k++) {
            float x = -(float)Num<T>::sub(lam[(size_t)g.col_idx[b + k] * Bp + f], ne[ne_at(b + k, f, g.E)]);
            sgn ^= __float_as_uint(x);
            c[k] = spa_compl(x);
            pre[k] = run;
            run = OpCompl::f(run, c[k]);
            ne[ne_at(b + k, f, g.E)] = (T)x;           // keep x for the sign of the edge
        }
        sgn = ~sgn;
        float suf = 0.0f;
        for (int k = d - 1; k >= 0; k--) {
            float others = OpCompl::f(pre[k], suf);
            float x = (float)ne[ne_at(b + k, f, g.E)];
            float v = __uint_as_float(__float_as_uint(spa_mag(others)) | ((sgn ^ __float_as_uint(x)) & 0x80000000u));
            ne[ne_at(b + k, f, g.E)] = (T)v;
            suf = OpCompl::f(suf, c[k]);
        }
    } else if (ALGO == LDPC_B200_ALGO_SPA_CM) {
        // Fast/CachedMult.hs:248-261: the row's StableDiv product (worst term held out, CachedMult.hs:25-56), folded from
        // the right in ascending column order, divided by the edge's own tanh.
        if (d <= 0) return;
        T t[MAXD];
        for (int k = 0; k < d; k++) {
            T v = Num<T>::sub(lam[(size_t)g.col_idx[b + k] * Bp + f], ne[ne_at(b + k, f, g.E)]);
            t[k] = Num<T>::tanh_(-(v / (T)2));
        }
        T wa, wb;                                                       // lit x | x >= 1 = (1, x) | otherwise = (x, 1)
        if (t[d - 1] >= (T)1) { wa = (T)1; wb = t[d - 1]; } else { wa = t[d - 1]; wb = (T)1; }
        for (int k = d - 2; k >= 0; k--) {                              // smult (lit t[k]) acc
            T la, lb;
            if (t[k] >= (T)1) { la = (T)1; lb = t[k]; } else { la = t[k]; lb = (T)1; }
            T mn, mx;
            if (Num<T>::abs_(la) < Num<T>::abs_(wa)) { mn = la; mx = wa; } else { mn = wa; mx = la; }
            wb = Num<T>::mul(Num<T>::mul(lb, mx), wb);
            wa = mn;
        }
        for (int k = 0; k < d; k++) {                                   // sdiv (a,b) c | a == c = b | otherwise = a * (b / c)
            T q = (wa == t[k]) ? wb : Num<T>::mul(wa, wb / t[k]);
            ne[ne_at(b + k, f, g.E)] = (T)-2 * atanh_prime<T>(q);
        }
    } else if (ALGO == LDPC_B200_ALGO_SPA) {
        // Orig.hs:86-91: tanh(-((lam_j - ne[m,j]) / 2)) per edge, leave-one-out LEFT fold of (*) from 1.
        T t[MAXD];
        for (int k = 0; k < d; k++) {
            T v = Num<T>::sub(lam[(size_t)g.col_idx[b + k] * Bp + f], ne[ne_at(b + k, f, g.E)]);
            t[k] = Num<T>::tanh_(-(v / (T)2));
        }
        for (int k = 0; k < d; k++) {
            T p = (T)1;
            for (int q = 0; q < d; q++) if (q != k) p = Num<T>::mul(p, t[q]);
            ne[ne_at(b + k, f, g.E)] = (T)-2 * atanh_prime<T>(p);
        }
    } else {
        // Min.hs:79-89 in two-minimum form (Model.hs:326-343): exact, because min and sign products are exact.
        T m1 = Num<T>::inf(), m2 = Num<T>::inf();
        int neg = 0;
        for (int k = 0; k < d; k++) {
            T x = -Num<T>::sub(lam[(size_t)g.col_idx[b + k] * Bp + f], ne[ne_at(b + k, f, g.E)]);
            T a = Num<T>::abs_(x);
            if (a < m1) { m2 = m1; m1 = a; } else if (a < m2) m2 = a;
            neg ^= (x < (T)0);
        }
        for (int k = 0; k < d; k++) {
            T x = -Num<T>::sub(lam[(size_t)g.col_idx[b + k] * Bp + f], ne[ne_at(b + k, f, g.E)]);
            T mag = (Num<T>::abs_(x) == m1) ? m2 : m1;
            T v = (neg ^ (x < (T)0)) ? -mag : mag;
            ne[ne_at(b + k, f, g.E)] = Num<T>::mul((T)-0.75, v);
        }
    }
}

// Orig.hs:95-98: lam'[j] = ne'[m1] + (ne'[m2] + (... + (ne'[mk] + L_j))). One thread per (variable, frame).
// ARRAYLET: the association of the arraylet decoders instead, L_j + (ne'[m1] + (ne'[m2] + ... + ne'[mk]))
// (Fast/CachedMult.hs:263-264 = Fast/Arraylet.hs:185-186).
template <typename T, bool ARRAYLET>
__global__ void k_vn(DeviceGraph g, const T *__restrict__ L, const T *__restrict__ ne, T *__restrict__ lam,
                     const uint8_t *__restrict__ state, int F, int Bp, int skip_done)
{
    int f = blockIdx.y * TPB + threadIdx.x, j = blockIdx.x;
    if (f >= F) return;
    if (skip_done && state[f]) return;
    const int q0 = g.col_ptr[j], q1 = g.col_ptr[j + 1];
    if (ARRAYLET) {
        if (q1 == q0) { lam[(size_t)j * Bp + f] = L[(size_t)j * Bp + f]; return; }
        T acc = ne[ne_at(g.col_edge[q1 - 1], f, g.E)];
        for (int q = q1 - 2; q >= q0; q--) acc = Num<T>::add(ne[ne_at(g.col_edge[q], f, g.E)], acc);
        lam[(size_t)j * Bp + f] = Num<T>::add(L[(size_t)j * Bp + f], acc);
    } else {
        T acc = L[(size_t)j * Bp + f];
        for (int q = q1 - 1; q >= q0; q--) acc = Num<T>::add(ne[ne_at(g.col_edge[q], f, g.E)], acc);
        lam[(size_t)j * Bp + f] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// TMA-staged kernels.  One CTA = TPB frames x CPC checks (VPC variables).  Every row a check touches is a contiguous
// TPB*sizeof(T)-byte segment of the frame-minor arrays, so the whole working set of the CTA is fetched with 1-D bulk
// copies (cp.async.bulk -> UBLKCP) signalled on one mbarrier, computed on from shared memory, and written back with
// bulk stores.  Why: with per-thread loads the compiler interleaves uses with later loads (predicate and register
// pressure heuristics), the in-order issue then stalls on a DRAM round trip several times per row and the kernel ran
// at 2.7 TB/s with 56 % long-scoreboard stalls; with bulk copies all bytes of the CTA are in flight at once and no
// register holds an address or a pending value.
// ------------------------------------------------------------------------------------------------------------------
// One check row of one frame from / to the staged rows in shared memory; returns the syndrome parity of the row.
// FULL: the row has exactly MAXD edges (no masks); otherwise it is padded to MAXD with the neutral element of the check
// operator, so both forms use the same predicate-free leave-one-out tree as the RESIDENT engine (~2.6 operations per edge).
template <typename T, int ALGO, int MAXD, bool FULL>
__device__ __forceinline__ int cn_row(T (*lam_r)[TPB], T (*ne_r)[TPB], int tid, int d)
{
    T x[MAXD];
    int par = 0;
#pragma unroll
    for (int k = 0; k < MAXD; k++) {
        const T l = lam_r[k][tid], n = ne_r[k][tid];                        // rows >= d: stale shared memory, masked below
        if (FULL) {
            par ^= (l > (T)0);                                              // syndrome of the CURRENT posteriors (Orig.hs:73-78)
            x[k] = -Num<T>::sub(l, n);                                      // x_j = -(lam_j - ne[m,j])   (Min.hs:80 / Orig.hs:84)
        } else {
            par ^= (k < d) & (l > (T)0);
            x[k] = (k < d) ? -Num<T>::sub(l, n) : Num<T>::inf();
        }
    }
    cn_row_core<T, ALGO, MAXD>(x, [&](auto ki, T v) {
        constexpr int k = decltype(ki)::value;
        if (FULL || k < d) ne_r[k][tid] = v;
    });
    return par;
}

// Check-node update: CTA = CPC consecutive checks x TPB frames; thread = one frame, loops over the CPC checks.
template <typename T, int ALGO, int MAXD, int CPC>
__global__ void __launch_bounds__(TPB) k_cn_tma(DeviceGraph g, const T *__restrict__ lam, T *__restrict__ ne,
                                                const uint8_t *__restrict__ state, uint8_t *__restrict__ unsat,
                                                int F, int Bp, int skip_done)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int S = MAXD + 2;
    constexpr uint32_t ROW = TPB * sizeof(T);
    T(*lam_s)[TPB] = reinterpret_cast<T(*)[TPB]>(smem_raw);                 // [CPC*MAXD][TPB]
    T(*ne_s)[TPB] = lam_s + CPC * MAXD;                                     // [CPC*MAXD][TPB]
    __shared__ int32_t s_rec[CPC * S];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, f0 = blockIdx.y * TPB, f = f0 + tid, m0 = blockIdx.x * CPC;
    const int nchk = min(CPC, g.M - m0);
    if (tid < nchk * S) s_rec[tid] = g.row_ell[(size_t)m0 * S + tid];       // one cooperative load: {b, d, columns} per check
    if (tid == 0) mbar_init(&bar, 1);
    const bool idle = f >= F || (skip_done && state[f]);
    if (!__syncthreads_or(!idle)) return;                                   // every frame of the block has stopped
    // producers: the lanes of warp 0 fetch the posterior rows, the lanes of warp 1 the message rows -- lane q owns copy q,
    // so the address arithmetic runs across the lanes and only the (warp-level) copy instruction itself is serialised.
    // (Spreading the copies over all four warps was measured 25 % slower.)
    if (tid == 0) {
        uint32_t rows = 0;
        for (int c = 0; c < nchk; c++) rows += (uint32_t)s_rec[c * S + 1];
        mbar_expect_tx(&bar, 2u * rows * ROW);
    }
    if (tid < 32) {
        for (int i = tid; i < nchk * MAXD; i += 32) {
            const int c = i / MAXD, k = i % MAXD;
            if (k < s_rec[c * S + 1]) bulk_g2s(lam_s[i], lam + (size_t)s_rec[c * S + 2 + k] * Bp + f0, ROW, &bar);
        }
    } else if (tid < 32 + nchk) {
        const int c = tid - 32;                                             // the check's message rows: one contiguous chunk
        if (s_rec[c * S + 1] > 0)
            bulk_g2s(ne_s[c * MAXD], ne + ((size_t)blockIdx.y * g.E + s_rec[c * S]) * TPB, (uint32_t)s_rec[c * S + 1] * ROW, &bar);
    }
    mbar_wait(&bar, 0);
    int par_any = 0;
    for (int c = 0; c < nchk; c++) {
        const int d = s_rec[c * S + 1];
        // rows of the full class degree (all rows of a regular code, the degree-18 rows of the AR4JA codes) take the
        // mask-free instance; the others are padded to MAXD with the neutral element
        if (d == MAXD) par_any |= cn_row<T, ALGO, MAXD, true>(lam_s + c * MAXD, ne_s + c * MAXD, tid, d);
        else par_any |= cn_row<T, ALGO, MAXD, false>(lam_s + c * MAXD, ne_s + c * MAXD, tid, d);
    }
    if (par_any && !idle) unsat[f] = 1;                                     // k_status reads it after this kernel
    // Frames of this block that have already stopped get their (never used again) messages overwritten too.
    fence_async_smem();
    __syncthreads();
    if (tid < nchk && s_rec[tid * S + 1] > 0) {
        bulk_s2g(ne + ((size_t)blockIdx.y * g.E + s_rec[tid * S]) * TPB, ne_s[tid * MAXD], (uint32_t)s_rec[tid * S + 1] * ROW);
        bulk_commit_wait_read();
    }
}

// Variable-node update (Orig.hs:95-98): CTA = VPC consecutive variables x TPB frames.  Rows per variable: its channel
// LLR and its (<= COL_ELL_D) messages in the order of the reference's right fold; the new posterior row goes back with a
// bulk store.
template <typename T, int VPC>
__global__ void __launch_bounds__(TPB) k_vn_tma(DeviceGraph g, const T *__restrict__ L, const T *__restrict__ ne,
                                                T *__restrict__ lam, const uint8_t *__restrict__ state,
                                                int F, int Bp, int skip_done)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int D = DeviceGraph::COL_ELL_D, S = D + 1, R = D + 1;        // R rows per variable: [0] = L, [1..] = messages
    constexpr uint32_t ROW = TPB * sizeof(T);
    T(*rows)[TPB] = reinterpret_cast<T(*)[TPB]>(smem_raw);                  // [VPC*R][TPB]
    __shared__ int32_t s_rec[VPC * S];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, f0 = blockIdx.y * TPB, f = f0 + tid, j0 = blockIdx.x * VPC;
    const int nvar = min(VPC, g.N - j0);
    if (tid < nvar * S) s_rec[tid] = g.col_ell[(size_t)j0 * S + tid];
    if (tid == 0) mbar_init(&bar, 1);
    const bool idle = f >= F || (skip_done && state[f]);
    if (!__syncthreads_or(!idle)) return;
    if (tid == 0) {
        uint32_t n = 0;
        for (int v = 0; v < nvar; v++) n += 1u + (uint32_t)s_rec[v * S];
        mbar_expect_tx(&bar, n * ROW);
    }
    if (tid < 64) {                                                         // lane q of warps 0 / 1 owns copy q (see k_cn_tma)
        for (int i = tid; i < nvar * R; i += 64) {
            const int v = i / R, r = i % R;
            if (r == 0) bulk_g2s(rows[i], L + (size_t)(j0 + v) * Bp + f0, ROW, &bar);
            else if (r - 1 < s_rec[v * S]) bulk_g2s(rows[i], ne + ((size_t)blockIdx.y * g.E + s_rec[v * S + r]) * TPB, ROW, &bar);
        }
    }
    mbar_wait(&bar, 0);
    for (int v = 0; v < nvar; v++) {
        const int d = s_rec[v * S];
        T acc = rows[v * R][tid];
#pragma unroll
        for (int i = 0; i < D; i++) if (i < d) acc = Num<T>::add(rows[v * R + 1 + i][tid], acc);   // last check first
        rows[v * R][tid] = acc;
    }
    // (frames of this block that have already stopped get a posterior they never read again)
    fence_async_smem();
    __syncthreads();
    if (tid < nvar) {
        bulk_s2g(lam + (size_t)(j0 + tid) * Bp + f0, rows[tid * R], ROW);
        bulk_commit_wait_read();
    }
}

__global__ void k_finish(const uint8_t *__restrict__ state, const int32_t *__restrict__ iters,
                         int32_t *__restrict__ out_iters, uint8_t *__restrict__ out_conv, int F)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    if (out_iters) out_iters[f] = iters[f];
    if (out_conv) out_conv[f] = state[f] == 1;
}

static int stream_alloc_buffers(ldpc_b200_decoder *d);

int stream_alloc(ldpc_b200_decoder *d)
{
    if (d->sb.active) return 0;            // `active` is the LAST buffer allocated: set only when all of them exist
    stream_free(d);
    int rc = stream_alloc_buffers(d);
    if (rc) stream_free(d);                // a partial allocation never survives a failed call
    return rc;
}

static int stream_alloc_buffers(ldpc_b200_decoder *d)
{
    StreamBuffers &s = d->sb;
    const ldpc_b200_code *c = d->code;
    int Bp = ((d->max_batch + TPB - 1) / TPB) * TPB;
    // Row stride of the frame-minor arrays.  All CTAs in flight work on the same block of frames, i.e. on the same offset
    // inside every row; with a power-of-two stride those addresses differ only in their high bits and pile onto a few
    // DRAM channels.  An odd number of TPB-frame blocks per row spreads consecutive rows over all channels.
#ifndef STREAM_NO_SKEW
    if (((Bp / TPB) & 1) == 0) Bp += TPB;
#endif
    size_t es = elem_size(d->precision);
    s.Bp = Bp;
    LDPC_CUDA(cudaMalloc(&s.L, (size_t)c->N * Bp * es));
    LDPC_CUDA(cudaMalloc(&s.lam, (size_t)c->N * Bp * es));
    LDPC_CUDA(cudaMalloc(&s.ne, (size_t)(c->E > 0 ? c->E : 1) * Bp * es));
    LDPC_CUDA(cudaMalloc(&s.iters, sizeof(int32_t) * Bp));
    LDPC_CUDA(cudaMalloc(&s.state, Bp));
    LDPC_CUDA(cudaMalloc(&s.unsat, Bp));
    LDPC_CUDA(cudaMalloc(&s.active, sizeof(int) * ACTIVE_SLOTS));        // one counter per iteration (zeroed once per decode)
    return 0;
}

void stream_free(ldpc_b200_decoder *d)
{
    StreamBuffers &s = d->sb;
    cudaFree(s.L); cudaFree(s.lam); cudaFree(s.ne); cudaFree(s.iters); cudaFree(s.state); cudaFree(s.unsat);
    cudaFree(s.active);
    s = StreamBuffers();
}

template <typename T, int ALGO, int MAXD, int CPC>
static int launch_cn_tma(ldpc_b200_decoder *d, int fb, int F, int skip_done)
{
    const DeviceGraph &g = d->code->g;
    StreamBuffers &s = d->sb;
    auto kern = k_cn_tma<T, ALGO, MAXD, CPC>;
    const size_t smem = (size_t)2 * CPC * MAXD * TPB * sizeof(T);
    if (smem > 48 * 1024) LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device
    kern<<<dim3((g.M + CPC - 1) / CPC, fb), TPB, smem, d->stream>>>(g, (const T *)s.lam, (T *)s.ne, s.state, s.unsat, F, s.Bp, skip_done);
    d->launches++;
    return 0;
}

template <typename T, int ALGO>
static int launch_cn(ldpc_b200_decoder *d, int fb, int F, int skip_done)
{
    const DeviceGraph &g = d->code->g;
    StreamBuffers &s = d->sb;
    // fp64 sum-product keeps the literal O(d^2) form; the cached-product variant is a verification statement too
    constexpr bool LITERAL64 = (ALGO == LDPC_B200_ALGO_SPA && sizeof(T) == 8) || ALGO == LDPC_B200_ALGO_SPA_CM;
    if constexpr (!LITERAL64) {
        switch (g.row_ell_d) {     // fixed-stride record class chosen at code creation (0: rows wider than 32)
        case 4: return launch_cn_tma<T, ALGO, 4, 8>(d, fb, F, skip_done);
        case 8: return launch_cn_tma<T, ALGO, 8, 4>(d, fb, F, skip_done);
        case 18: return launch_cn_tma<T, ALGO, 18, 2>(d, fb, F, skip_done);   // 2 checks per CTA measured best (1: 11.95 ms, 2: 11.27, 4: 12.47)
        case 32: return launch_cn_tma<T, ALGO, 32, 1>(d, fb, F, skip_done);
        default: break;
        }
    }
    dim3 grid(g.M, fb);
    if (g.max_row_deg <= 8)
        k_cn<T, ALGO, 8><<<grid, TPB, 0, d->stream>>>(g, (const T *)s.lam, (T *)s.ne, s.state, s.unsat, F, s.Bp, skip_done);
    else if (g.max_row_deg <= 32)
        k_cn<T, ALGO, 32><<<grid, TPB, 0, d->stream>>>(g, (const T *)s.lam, (T *)s.ne, s.state, s.unsat, F, s.Bp, skip_done);
    else
        k_cn<T, ALGO, 128><<<grid, TPB, 0, d->stream>>>(g, (const T *)s.lam, (T *)s.ne, s.state, s.unsat, F, s.Bp, skip_done);
    d->launches++;
    return 0;
}

template <typename T>
static int launch_vn(ldpc_b200_decoder *d, int fb, int F, int skip_done)
{
    const DeviceGraph &g = d->code->g;
    StreamBuffers &s = d->sb;
    if (d->algo >= LDPC_B200_ALGO_SPA_CM) {        // arraylet association: SPA_CM, SPA_ARRAYLET, MINSUM_ARRAYLET
        k_vn<T, true><<<dim3(g.N, fb), TPB, 0, d->stream>>>(g, (const T *)s.L, (const T *)s.ne, (T *)s.lam, s.state, F, s.Bp, skip_done);
    } else if (g.max_col_deg <= DeviceGraph::COL_ELL_D) {
        constexpr int VPC = 4;
        auto kern = k_vn_tma<T, VPC>;
        const size_t smem = (size_t)VPC * (DeviceGraph::COL_ELL_D + 1) * TPB * sizeof(T);
        if (smem > 48 * 1024) LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dim3((g.N + VPC - 1) / VPC, fb), TPB, smem, d->stream>>>(g, (const T *)s.L, (const T *)s.ne, (T *)s.lam, s.state, F, s.Bp, skip_done);
    } else {
        k_vn<T, false><<<dim3(g.N, fb), TPB, 0, d->stream>>>(g, (const T *)s.L, (const T *)s.ne, (T *)s.lam, s.state, F, s.Bp, skip_done);
    }
    d->launches++;
    return 0;
}

template <typename T>
static int stream_decode_t(ldpc_b200_decoder *d, int max_iters, int early_exit, int F, const void *d_llr,
                           bool llr_is_double, uint32_t *d_bits, int32_t *d_iters, uint8_t *d_conv,
                           double *h_c2v, double *h_lam)
{
    const ldpc_b200_code *c = d->code;
    const DeviceGraph &g = c->g;
    StreamBuffers &s = d->sb;
    cudaStream_t st = d->stream;
    int Bp = s.Bp, N = c->N, E = c->E, M = c->M;
    if (g.max_row_deg > 128) { set_error("STREAM engine supports check degree <= 128 (got %d)", g.max_row_deg); return LDPC_B200_EUNSUPPORTED; }
    int fb = (F + TPB - 1) / TPB;
    int skip = early_exit ? 1 : 0;

    LDPC_CUDA(cudaMemsetAsync(s.ne, 0, (size_t)(E > 0 ? E : 1) * Bp * sizeof(T), st));
    LDPC_CUDA(cudaMemsetAsync(s.state, 0, Bp, st));
    LDPC_CUDA(cudaMemsetAsync(s.unsat, 0, Bp, st));
    LDPC_CUDA(cudaMemsetAsync(s.iters, 0, sizeof(int32_t) * Bp, st));
    LDPC_CUDA(cudaMemsetAsync(s.active, 0, sizeof(int) * ACTIVE_SLOTS, st));
    {
        dim3 grid((N + 31) / 32, (Bp + 31) / 32), blk(32, 8);
        if (llr_is_double) k_load<T, double><<<grid, blk, 0, st>>>((const double *)d_llr, (T *)s.L, (T *)s.lam, N, F, Bp);
        else k_load<T, float><<<grid, blk, 0, st>>>((const float *)d_llr, (T *)s.L, (T *)s.lam, N, F, Bp);
        d->launches++;
    }
    // Grid layout: blockIdx.x = check / variable / word, blockIdx.y = block of 128 frames.  CTAs are dispatched x-fastest,
    // so all checks of one frame block run back to back and the posterior lines they gather (N x 512 B) stay in L2:
    // DRAM sees each posterior once per phase, which is what the algorithmic-bytes model assumes.
    // Per iteration n: k_cn (syndrome of lam_n + new messages) -> k_status -> k_hard -> k_vn.  A frame whose syndrome
    // turns out to be zero has computed one set of messages it never uses; in exchange there is no separate syndrome pass.
    std::vector<T> tmp;
    for (int n = 0;; n++) {
        dim3 gcn(M, fb);
        if (n >= max_iters) {
            if (M > 0) { k_syndrome<T><<<gcn, TPB, 0, st>>>(g, (const T *)s.lam, s.state, s.unsat, F, Bp, 1); d->launches++; }
        } else if (M > 0) {
            // converged frames are skipped via `state` when early exit is on; with early_exit = 0 they keep iterating
            // (the arraylet statements share the check updates of Arraylet.hs:176-183 = Orig.hs:81-92 and
            // ArrayletMin.hs:182-189 = Min.hs:75-89; only the posterior association differs, see launch_vn)
            const int a = d->algo;
            int rc = (a == LDPC_B200_ALGO_SPA || a == LDPC_B200_ALGO_SPA_ARRAYLET) ? launch_cn<T, LDPC_B200_ALGO_SPA>(d, fb, F, skip)
                     : a == LDPC_B200_ALGO_SPA_CM                                  ? launch_cn<T, LDPC_B200_ALGO_SPA_CM>(d, fb, F, skip)
                                                                                   : launch_cn<T, LDPC_B200_ALGO_MINSUM>(d, fb, F, skip);
            if (rc) return rc;
        }
        // frames still running after the test of iteration n are counted in slot n of `active` (all slots zeroed above)
        if (n > 0 && n % ACTIVE_SLOTS == 0) LDPC_CUDA(cudaMemsetAsync(s.active, 0, sizeof(int) * ACTIVE_SLOTS, st));
        k_status<<<(F + 255) / 256, 256, 0, st>>>(s.state, s.unsat, s.iters, s.active + n % ACTIVE_SLOTS, n, max_iters, F);
        d->launches++;
        k_hard<T><<<dim3((words_per_frame(N) + HARD_WPC - 1) / HARD_WPC, fb), TPB, 0, st>>>((const T *)s.lam, (const T *)s.L, s.state, s.iters, d_bits, n, N, F, Bp);
        d->launches++;
        if (n >= max_iters) break;
        // "Has every frame stopped?" costs a device -> host copy and a stream synchronisation, during which the GPU idles: asked
        // only every SYNC_EVERY iterations (and while tracing); the launches issued past the last active iteration find all their
        // frames idle and exit at once.
        if (early_exit && (n % SYNC_EVERY == SYNC_EVERY - 1 || n == 0 || h_c2v || h_lam)) {
            int active = 0;
            LDPC_CUDA(cudaMemcpyAsync(&active, s.active + n % ACTIVE_SLOTS, sizeof(int), cudaMemcpyDeviceToHost, st));
            LDPC_CUDA(cudaStreamSynchronize(st));
            if (active == 0) break;
        }
        { int rc = launch_vn<T>(d, fb, F, skip); if (rc) return rc; }
        if (h_c2v || h_lam) {   // trace of frame 0: column 0 of the frame-minor arrays
            size_t need = (size_t)(E > N ? E : N);
            tmp.resize(need);
            LDPC_CUDA(cudaStreamSynchronize(st));
            if (h_c2v && E > 0) {
                LDPC_CUDA(cudaMemcpy2D(tmp.data(), sizeof(T), s.ne, (size_t)TPB * sizeof(T), sizeof(T), E, cudaMemcpyDeviceToHost));   // frame 0: block 0, lane 0
                for (int e = 0; e < E; e++) h_c2v[(size_t)n * E + e] = (double)tmp[e];
            }
            if (h_lam) {
                LDPC_CUDA(cudaMemcpy2D(tmp.data(), sizeof(T), s.lam, (size_t)Bp * sizeof(T), sizeof(T), N, cudaMemcpyDeviceToHost));
                for (int j = 0; j < N; j++) h_lam[(size_t)n * N + j] = (double)tmp[j];
            }
        }
    }
    k_finish<<<(F + 255) / 256, 256, 0, st>>>(s.state, s.iters, d_iters, d_conv, F);
    d->launches++;
    LDPC_CUDA(cudaGetLastError());
    return 0;
}

int stream_decode(ldpc_b200_decoder *d, int max_iters, int early_exit, int n_frames, const void *d_llr,
                  bool llr_is_double, uint32_t *d_bits_packed, int32_t *d_iters, uint8_t *d_conv,
                  double *h_trace_c2v, double *h_trace_lam)
{
    int rc = stream_alloc(d);
    if (rc) return rc;
    if (n_frames > d->max_batch) { set_error("n_frames %d exceeds decoder max_batch %d", n_frames, d->max_batch); return LDPC_B200_EINVAL; }
    if (d->precision == LDPC_B200_PREC_FP64)
        return stream_decode_t<double>(d, max_iters, early_exit, n_frames, d_llr, llr_is_double, d_bits_packed, d_iters, d_conv, h_trace_c2v, h_trace_lam);
    return stream_decode_t<float>(d, max_iters, early_exit, n_frames, d_llr, llr_is_double, d_bits_packed, d_iters, d_conv, h_trace_c2v, h_trace_lam);
}

}  // namespace ldpc
