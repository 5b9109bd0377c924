// onchip_engine.cu -- ONCHIP engine: flooding BP for ANY parity-check matrix whose per-frame state fits shared memory
// (moon.7.13, 1920.1280.3.303, user codes), whole frame on chip for all iterations.
//
// The RESIDENT kernels are compiled per quasi-cyclic code (circulant tables as immediates).  This is the run-time
// table counterpart: same frame life cycle (persistent CTAs, atomic work counter, per-frame early exit, fused Philox
// generation and error counting for ldpc_b200_simulate), but the graph comes from the fixed-stride records of
// DeviceGraph (row_soa / col_soa, transposed so that a warp reads one line per record entry), read through L1, and all three state arrays live in shared memory:
//     lam[N]  posteriors      L[N]  channel LLRs      ne[M][DP]  check->variable messages, DP = row class + 1 (odd, so
//                                                     that the threads of consecutive checks hit distinct banks)
// One frame = TPF threads (32..256): thread t owns checks t, t+TPF, ... in the check phase and variables t, t+TPF, ...
// in the variable phase.  Reference lines: driver loop Orig.hs:58-71; syndrome Orig.hs:73-78; check update Min.hs:75-89 /
// Orig.hs:81-92 (fp32 sum-product in the complement-domain tree of bp_tree.cuh); variable update Orig.hs:95-98 with
// the reference's right-fold association, so the float64 instance reproduces the reference arithmetic bit for bit and the float32
// instance to the STREAM engine.
#include "common.cuh"
#include "philox.cuh"
#include <algorithm>
#include <utility>

namespace ldpc {

#include "bp_tree.cuh"
#include "resident_common.cuh"

struct OnchipGeom {
    int TPF, FPC, DP, CD;     // threads per frame, frames per CTA, message row stride, staged entries per variable record
    size_t frame_elems;       // 2N + M*DP
    size_t table_bytes;       // graph records staged in shared memory (uint16), multiple of 16
};

template <typename T, int ALGO, int MAXD, bool REG>
__global__ void __launch_bounds__(512) k_onchip(DeviceGraph g, ResidentParams p, OnchipGeom geo)
{
    typedef FB<T> F;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = g.M, N = g.N, W = (N + 31) / 32, TPF = geo.TPF, DP = geo.DP, CD = geo.CD;
    const int slot = threadIdx.x / TPF, tf = threadIdx.x % TPF, lane = threadIdx.x & 31;
    const int bar = slot + 1, WPF = TPF / 32, wif = tf >> 5;
    // graph records, staged once per (persistent) CTA and shared by all its frames: 16-bit, structure of arrays
    //   rowT[(MAXD+1)][M] = {degree, columns...}      colT[(CD+1)][N] = {degree, message slots in fold order...}
    uint16_t *rowT = reinterpret_cast<uint16_t *>(smem_raw);
    uint16_t *colT = rowT + (size_t)(MAXD + 1) * M;
    for (int i = threadIdx.x; i < (MAXD + 1) * M; i += blockDim.x) rowT[i] = (uint16_t)g.row_soa[i];
    for (int i = threadIdx.x; i < (CD + 1) * N; i += blockDim.x) colT[i] = (uint16_t)g.col_soa[i];
    T *lam = reinterpret_cast<T *>(smem_raw + geo.table_bytes) + (size_t)slot * geo.frame_elems;
    T *L = lam + N;
    T *ne = L + N;
    unsigned int *s_next = reinterpret_cast<unsigned int *>(reinterpret_cast<T *>(smem_raw + geo.table_bytes) + (size_t)geo.FPC * geo.frame_elems);
    __syncthreads();

    FrameCounters cnt;
    for (;;) {
        if (tf == 0) s_next[slot] = atomicAdd(p.work, 1u);
        bar_sync(bar, TPF);
        const unsigned int f = s_next[slot];
        if (f >= p.n_frames) break;

        // ---- frame in: lam = L (Orig.hs:58), ne = 0 (Orig.hs:64-65) ----
        if (p.llr) {
            const T *src = reinterpret_cast<const T *>(p.llr) + (size_t)f * N;
            for (int j = tf; j < N; j += TPF) { T v = src[j]; L[j] = v; lam[j] = v; }
        } else {
            for (int q4 = tf; q4 < (N + 3) / 4; q4 += TPF) {
                float nz[4];
                awgn4(p.seed, p.first_frame + f, (uint32_t)q4, nz);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    int j = 4 * q4 + q;
                    if (j < N) { T v = (j < p.n_tx) ? (T)awgn_llr(nz[q], 0, p.sigma, p.scale) : (T)0; L[j] = v; lam[j] = v; }
                }
            }
        }
        for (int e = tf; e < M * DP; e += TPF) ne[e] = (T)0;
        bar_sync(bar, TPF);

        bool done = false;
        for (int n = 0;; n++) {
            bool finished_now = false, conv = false;
            if (!(done && n == p.max_iters)) {
                // ---- check phase (+ syndrome of the current posteriors) ----
                int unsat = 0;
                for (int m = tf; m < M; m += TPF) {
                    const int d = REG ? MAXD : rowT[m];                      // REG: every check has exactly MAXD edges
                    T *row = ne + (size_t)m * DP;
                    T x[MAXD];
                    int par = 0;
#pragma unroll
                    for (int k = 0; k < MAXD; k++) {
                        const T l = lam[rowT[(1 + k) * M + m]], v = row[k];     // padding repeats the last column; masked
                        if (REG) {
                            par ^= (l > (T)0);                               // Orig.hs:73-78
                            x[k] = -F::sub(l, v);                            // x_j = -(lam_j - ne[m,j])
                        } else {
                            par ^= (k < d) & (l > (T)0);
                            x[k] = (k < d) ? -F::sub(l, v) : F::inf();
                        }
                    }
                    unsat |= par;
                    if (n < p.max_iters)
                        cn_row_core<T, ALGO, MAXD>(x, [&](auto ki, T v) {
                            constexpr int k = decltype(ki)::value;
                            if (REG || k < d) row[k] = v;
                        });
                }
                const bool any = bar_or(bar, TPF, unsat != 0);               // also: messages visible to the variable phase
                if (!done && (!any || n == p.max_iters)) { finished_now = true; conv = !any; }
            }
            if (finished_now) {
                // syndrome ok -> hard(lam); iterations exhausted -> hard(orig_lam)   (Orig.hs:68-71)
                done = true;
                const T *src = conv ? lam : L;
                unsigned int errs = 0;
                for (int w = wif; w < W; w += WPF) {
                    const int j = 32 * w + lane;
                    unsigned int word = __ballot_sync(0xffffffffu, j < N && src[j] > (T)0);
                    if (lane == 0) {
                        if (p.bits) p.bits[(size_t)f * W + w] = word;
                        int lo = 32 * w;
                        if (lo + 32 <= p.k_info) errs += __popc(word);
                        else if (lo < p.k_info) errs += __popc(word & ((1u << (p.k_info - lo)) - 1u));
                    }
                }
                if (tf == 0) {
                    if (p.iters) p.iters[f] = n;
                    if (p.conv) p.conv[f] = conv ? 1 : 0;
                }
                if (p.counters) {
                    bool anyerr = bar_or(bar, TPF, errs != 0);
                    cnt.berr += errs;
                    if (tf == 0) { cnt.frames++; cnt.ferr += anyerr; cnt.iters += n; cnt.unconv += !conv; }
                } else if (!p.early_exit) {
                    bar_sync(bar, TPF);          // decisions read before the variable phase rewrites lam
                }
            }
            if (n == p.max_iters || (done && p.early_exit)) break;

            // ---- variable phase (Orig.hs:95-98): lam'[j] = ne'[m1] + (ne'[m2] + (... + (ne'[mk] + L_j))) ----
            for (int j = tf; j < N; j += TPF) {
                const int d = colT[j];
                T acc = L[j];
#pragma unroll
                for (int i = 0; i < DeviceGraph::COL_ELL_D; i++)
                    if (i < CD && i < d) acc = F::add(ne[colT[(1 + i) * N + j]], acc);                           // last check first
                if (d > CD) {
                    const int q0 = g.col_ptr[j];
                    for (int q = q0 + d - 1 - CD; q >= q0; q--) acc = F::add(ne[g.col_slot[q]], acc);
                }
                lam[j] = acc;
            }
            bar_sync(bar, TPF);
        }
    }

    if (p.counters) {
        if (lane == 0 && cnt.berr) atomicAdd(&p.counters[2], cnt.berr);
        if (tf == 0 && cnt.frames) {
            atomicAdd(&p.counters[0], cnt.frames); atomicAdd(&p.counters[1], cnt.ferr);
            atomicAdd(&p.counters[3], cnt.iters); atomicAdd(&p.counters[4], cnt.unconv);
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------

static const size_t ONCHIP_MAX_CTA_BYTES = 113 * 1024;      // tables + one frame: at least two CTAs per SM

static bool onchip_geometry(const ldpc_b200_code *c, size_t es, OnchipGeom *out)
{
    const DeviceGraph &g = c->g;
    if (!g.row_ell_d || !g.row_soa || !g.col_soa || c->M <= 0 || c->E <= 0) return false;
    OnchipGeom geo;
    geo.DP = g.row_ell_d + 1;
    geo.CD = std::min(c->max_col_deg, (int)DeviceGraph::COL_ELL_D);
    if (c->N > 65535 || (size_t)c->M * geo.DP > 65535) return false;                 // 16-bit records
    geo.frame_elems = (size_t)2 * c->N + (size_t)c->M * geo.DP;
    geo.table_bytes = (((size_t)(g.row_ell_d + 1) * c->M + (size_t)(geo.CD + 1) * c->N) * sizeof(uint16_t) + 15) / 16 * 16;
    const size_t frame_bytes = geo.frame_elems * es;
    if (geo.table_bytes + frame_bytes > ONCHIP_MAX_CTA_BYTES) return false;
    int work = std::max(c->M, c->N);
    geo.TPF = std::min(256, std::max(32, ((work + 5) / 6 + 31) / 32 * 32));
    geo.FPC = (int)std::min<size_t>(std::min(8, 512 / geo.TPF), (ONCHIP_MAX_CTA_BYTES - geo.table_bytes) / frame_bytes);
    if (geo.FPC < 1) geo.FPC = 1;
    if (out) *out = geo;
    return true;
}

bool onchip_fits(const ldpc_b200_code *c, int precision) { return onchip_geometry(c, elem_size(precision), nullptr); }

template <typename T, int ALGO, int MAXD, bool REG>
static int onchip_launch_t(ldpc_b200_decoder *d, ResidentParams &p, const OnchipGeom &geo)
{
    auto kern = k_onchip<T, ALGO, MAXD, REG>;
    const int threads = geo.FPC * geo.TPF;
    const size_t smem = geo.table_bytes + (size_t)geo.FPC * geo.frame_elems * sizeof(T) + geo.FPC * sizeof(unsigned int);
    int blocks_per_sm = 0, n_sm = 0;
    LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    LDPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, threads, smem));
    LDPC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, d->code->device));
    if (blocks_per_sm < 1) { set_error("ONCHIP kernel does not fit on this device"); return LDPC_B200_EUNSUPPORTED; }
    if (!d->res_work) LDPC_CUDA(cudaMalloc(&d->res_work, sizeof(unsigned int)));
    LDPC_CUDA(cudaMemsetAsync(d->res_work, 0, sizeof(unsigned int), d->stream));
    p.work = d->res_work;
    unsigned int groups = (p.n_frames + geo.FPC - 1) / geo.FPC;
    unsigned int grid = std::min<unsigned int>(groups, (unsigned int)(blocks_per_sm * n_sm));
    if (grid == 0) return 0;
    kern<<<grid, threads, smem, d->stream>>>(d->code->g, p, geo);
    d->launches++;
    LDPC_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, int ALGO>
static int onchip_launch_d(ldpc_b200_decoder *d, ResidentParams &p, const OnchipGeom &geo)
{
    const bool reg = d->code->g.row_regular != 0;
    switch (d->code->g.row_ell_d) {
    case 4: return reg ? onchip_launch_t<T, ALGO, 4, true>(d, p, geo) : onchip_launch_t<T, ALGO, 4, false>(d, p, geo);
    case 8: return reg ? onchip_launch_t<T, ALGO, 8, true>(d, p, geo) : onchip_launch_t<T, ALGO, 8, false>(d, p, geo);
    case 18: return reg ? onchip_launch_t<T, ALGO, 18, true>(d, p, geo) : onchip_launch_t<T, ALGO, 18, false>(d, p, geo);
    case 32: return reg ? onchip_launch_t<T, ALGO, 32, true>(d, p, geo) : onchip_launch_t<T, ALGO, 32, false>(d, p, geo);
    default: set_error("ONCHIP engine: check degree above 32"); return LDPC_B200_EUNSUPPORTED;
    }
}

int onchip_launch(ldpc_b200_decoder *d, ResidentParams &p)
{
    OnchipGeom geo;
    if (!onchip_geometry(d->code, elem_size(d->precision), &geo)) { set_error("ONCHIP engine: the frame state of this code does not fit shared memory"); return LDPC_B200_EUNSUPPORTED; }
    if (d->algo == LDPC_B200_ALGO_SPA) {
        if (d->precision != LDPC_B200_PREC_FP32) { set_error("ONCHIP engine: sum-product is float32 only"); return LDPC_B200_EUNSUPPORTED; }
        return onchip_launch_d<float, LDPC_B200_ALGO_SPA>(d, p, geo);
    }
    if (d->algo != LDPC_B200_ALGO_MINSUM) { set_error("ONCHIP engine: unsupported algorithm"); return LDPC_B200_EUNSUPPORTED; }
    return d->precision == LDPC_B200_PREC_FP64 ? onchip_launch_d<double, LDPC_B200_ALGO_MINSUM>(d, p, geo)
                                               : onchip_launch_d<float, LDPC_B200_ALGO_MINSUM>(d, p, geo);
}

}  // namespace ldpc
