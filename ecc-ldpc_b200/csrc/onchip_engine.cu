// onchip_engine.cu -- ONCHIP engine: flooding BP for ANY parity-check matrix whose per-frame state fits shared memory
// (moon.7.13, 1920.1280.3.303, user codes), whole frame on chip for all iterations.
//
// The RESIDENT kernels are compiled per quasi-cyclic code (circulant tables as immediates).  This is the run-time
// table counterpart: same frame life cycle (persistent CTAs, atomic work counter, per-frame early exit, fused Philox
// generation and error counting for ldpc_b200_simulate), but the graph comes from 16-bit records staged once per
// persistent CTA into shared memory, next to the state of every frame of the CTA:
//     lam[N]  posteriors      L[N]  channel LLRs      ne[SE]  check->variable messages
// One frame = TPF threads (32..256): thread t owns checks t, t+TPF, ... in the check phase and variable POSITIONS
// t, t+TPF, ... in the variable phase.  The kernel is bound by shared-memory bandwidth, and with the natural placement
// a quarter of its wavefronts were bank conflicts of the two gathers (posteriors by the checks, messages by the
// variables).  onchip_build_tables therefore chooses the placement:
//   * message slots: every edge lies in one "check group" (32 consecutive checks, edge position k) and one "variable
//     group" (32 consecutive positions, fold position i); a proper 32-colouring of the edges of that bipartite
//     multigraph (Koenig: it exists, maximum degree 32) is a bank assignment under which BOTH phases touch every bank
//     at most once per warp access; slot = bank + 32 * rank;
//   * variable positions: a permutation inside every block of 32 variables (so the variable phase stays conflict-free
//     by construction), chosen greedily to spread the variables of every check group over distinct banks.
// Reference lines: driver loop Orig.hs:58-71; syndrome Orig.hs:73-78; check update Min.hs:75-89 / Orig.hs:81-92 (fp32
// sum-product in the complement-domain tree of bp_tree.cuh); variable update Orig.hs:95-98 with the reference's
// right-fold association, so the float64 instance reproduces the reference arithmetic bit for bit and the float32
// instance equals the STREAM engine.
#include "common.cuh"
#include "philox.cuh"
#include <algorithm>
#include <atomic>
#include <utility>

namespace ldpc {

#include "bp_tree.cuh"
#include "resident_common.cuh"

struct OnchipLaunch { cudaStream_t stream; bool static_grid; };

struct OnchipGeom {
    int TPF, FPC, SE, CD;     // threads per frame, frames per CTA, message slots per frame, staged entries per variable record
    size_t frame_elems;       // 2N + SE
    size_t table_bytes;       // graph records staged in shared memory (uint16), multiple of 16
};

template <typename T, int ALGO, int MAXD, bool REG, int CDR>
__global__ void __launch_bounds__(512) k_onchip(DeviceGraph g, ResidentParams p, OnchipGeom geo)
{
    typedef FB<T> F;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = g.M, N = g.N, W = (N + 31) / 32, TPF = geo.TPF, SE = geo.SE;
    const int slot = threadIdx.x / TPF, tf = threadIdx.x % TPF, lane = threadIdx.x & 31;
    const int bar = slot + 1, WPF = TPF / 32, wif = tf >> 5;
    // graph records (see onchip_build_tables), staged once per persistent CTA and shared by all its frames
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(g.oc_tab);
        uint4 *dst = reinterpret_cast<uint4 *>(smem_raw);
        for (int i = threadIdx.x; i < g.oc_tab_elems / 8; i += blockDim.x) dst[i] = src[i];
    }
    constexpr int REC4 = (2 * MAXD + 7) / 8;
    constexpr int ES = sizeof(T) / 4;                        // offsets are stored for 4-byte elements
    const uint4 *rowA = reinterpret_cast<const uint4 *>(smem_raw);                              // [REC4][M]
    const uint16_t *rowDeg = reinterpret_cast<const uint16_t *>(rowA + (size_t)REC4 * M);       // [M], irregular codes only
    const uint16_t *colA = rowDeg + (REG ? 0 : (M + 7) / 8 * 8);                                // [N][CDR]
    const uint16_t *varPos = colA + ((size_t)N * CDR + 7) / 8 * 8;                              // [N]
    T *lam = reinterpret_cast<T *>(smem_raw + geo.table_bytes) + (size_t)slot * geo.frame_elems;
    T *L = lam + N;
    T *ne = L + N;
    unsigned int *s_next = reinterpret_cast<unsigned int *>(reinterpret_cast<T *>(smem_raw + geo.table_bytes) + (size_t)geo.FPC * geo.frame_elems);
    __syncthreads();

    FrameCounters cnt;
    for (int round = 0;; round++) {
        unsigned int f;
        if (!p.work) {                       // static assignment (small latency-bound launches, see resident_kernel.cuh)
            if (round) break;
            f = blockIdx.x * geo.FPC + slot;
        } else {
            if (tf == 0) s_next[slot] = atomicAdd(p.work, 1u);
            bar_sync(bar, TPF);
            f = s_next[slot];
        }
        if (f >= p.n_frames) break;
        const unsigned int fi = p.use_map ? p.slot_map[f] : f;      // output index: the frame itself or its mailbox slot

        // ---- frame in: lam = L (Orig.hs:58), ne = 0 (Orig.hs:64-65) ----
        if (p.llr) {
            const T *src = reinterpret_cast<const T *>(p.llr) + (size_t)fi * N;
            for (int j = tf; j < N; j += TPF) { T v = src[j]; int q = varPos[j]; L[q] = v; lam[q] = v; }
        } else if (p.llr64) {
            // the reference's float64 vector rounded to T; positions >= take are punctured (Utils.hs:55)
            const double *src = p.llr64 + (size_t)fi * p.llr64_stride;
            for (int j = tf; j < N; j += TPF) { T v = j < p.llr64_take ? (T)src[j] : (T)0; int q = varPos[j]; L[q] = v; lam[q] = v; }
        } else {
            for (int q4 = tf; q4 < (N + 3) / 4; q4 += TPF) {
                float nz[4];
                awgn4(p.seed, p.first_frame + f, (uint32_t)q4, nz);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    int j = 4 * q4 + q;
                    if (j < N) { T v = (j < p.n_tx) ? (T)awgn_llr(nz[q], 0, p.sigma, p.scale) : (T)0; int pos = varPos[j]; L[pos] = v; lam[pos] = v; }
                }
            }
        }
        for (int e = tf; e < SE; e += TPF) ne[e] = (T)0;
        bar_sync(bar, TPF);

        bool done = false;
        for (int n = 0;; n++) {
            bool finished_now = false, conv = false;
            if (!(done && n == p.max_iters)) {
                // ---- check phase (+ syndrome of the current posteriors) ----
                int unsat = 0;
                for (int m = tf; m < M; m += TPF) {
                    const int d = REG ? MAXD : rowDeg[m];                    // REG: every check has exactly MAXD edges
                    uint32_t rec[REC4 * 4];                                  // 2 offsets per word
#pragma unroll
                    for (int q = 0; q < REC4; q++) {
                        const uint4 v = rowA[q * M + m];
                        rec[4 * q] = v.x; rec[4 * q + 1] = v.y; rec[4 * q + 2] = v.z; rec[4 * q + 3] = v.w;
                    }
                    auto off = [&](int idx) -> uint32_t { return ((idx & 1) ? (rec[idx >> 1] >> 16) : (rec[idx >> 1] & 0xffffu)) * ES; };
                    T x[MAXD];
                    int par = 0;
#pragma unroll
                    for (int k = 0; k < MAXD; k++) {                         // padding entries repeat the last edge; masked
                        const T l = *reinterpret_cast<const T *>(reinterpret_cast<const char *>(lam) + off(k));
                        const T v = *reinterpret_cast<const T *>(reinterpret_cast<const char *>(ne) + off(MAXD + k));
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
                            if (REG || k < d) *reinterpret_cast<T *>(reinterpret_cast<char *>(ne) + off(MAXD + k)) = v;
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
                    unsigned int word = __ballot_sync(0xffffffffu, j < N && src[varPos[j < N ? j : 0]] > (T)0);
                    if (lane == 0) {
                        if (p.bits) p.bits[(size_t)fi * W + w] = word;
                        int lo = 32 * w;
                        if (lo + 32 <= p.k_info) errs += __popc(word);
                        else if (lo < p.k_info) errs += __popc(word & ((1u << (p.k_info - lo)) - 1u));
                    }
                }
                if (tf == 0) {
                    if (p.iters) p.iters[fi] = n;
                    if (p.conv) p.conv[fi] = conv ? 1 : 0;
                }
                if (p.done) {                    // mailbox: completion flag behind a system-scope fence (arrays in mapped host memory)
                    __threadfence_system();
                    bar_sync(bar, TPF);
                    if (tf == 0) *reinterpret_cast<volatile unsigned int *>(p.done + fi) = p.done_seq;
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
            if (p.trace_c2v && f == 0) {         // message tap (run-time flag: the messages live in shared memory, nothing else changes)
                T *out = reinterpret_cast<T *>(p.trace_c2v) + (size_t)n * p.trace_ne;
                for (int e = tf; e < SE; e += TPF) out[e] = ne[e];      // slot order; the host maps slots back to reference edges
            }

            // ---- variable phase (Orig.hs:95-98): lam'[j] = ne'[m1] + (ne'[m2] + (... + (ne'[mk] + L_j))), by position ----
            for (int q = tf; q < N; q += TPF) {
                uint32_t rec[CDR / 2];
                if constexpr (CDR == 4) { const uint2 v = reinterpret_cast<const uint2 *>(colA)[q]; rec[0] = v.x; rec[1] = v.y; }
                else { const uint4 v = reinterpret_cast<const uint4 *>(colA)[q]; rec[0] = v.x; rec[1] = v.y; rec[2] = v.z; rec[3] = v.w; }
                const int d = rec[0] & 0xffffu;
                T acc = L[q];
#pragma unroll
                for (int i = 0; i < CDR - 1; i++) {                                                              // last check first
                    const uint32_t o = (((i + 1) & 1) ? (rec[(i + 1) >> 1] >> 16) : (rec[(i + 1) >> 1] & 0xffffu)) * ES;
                    if (i < d) acc = F::add(*reinterpret_cast<const T *>(reinterpret_cast<const char *>(ne) + o), acc);
                }
                if (d > CDR - 1) {
                    const int q0 = g.col_ptr[g.oc_posvar[q]];
                    for (int e = q0 + d - CDR; e >= q0; e--) acc = F::add(ne[g.col_slot[e]], acc);
                }
                lam[q] = acc;
            }
            bar_sync(bar, TPF);
            if (p.trace_lam && f == 0) {
                T *out = reinterpret_cast<T *>(p.trace_lam) + (size_t)n * N;
                for (int q = tf; q < N; q += TPF) out[q] = lam[q];      // position order; reordered by the host
            }
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

static constexpr int ONCHIP_MAX_DEVICES = 64;
static const size_t ONCHIP_MAX_CTA_BYTES = 113 * 1024;      // tables + frames of a CTA when two CTAs share an SM

// Proper 32-colouring of the edges of a bipartite multigraph with maximum degree <= 32 (left node ga[e], right node
// gb[e]): the textbook alternating-path algorithm.  colour[e] in [0, 32).
static bool colour_edges_32(int nA, int nB, const std::vector<int> &ga, const std::vector<int> &gb, std::vector<int> &colour)
{
    const int E = (int)ga.size(), C = 32;
    std::vector<int> atA((size_t)nA * C, -1), atB((size_t)nB * C, -1);
    colour.assign(E, -1);
    for (int e = 0; e < E; e++) {
        const int u = ga[e], v = gb[e];
        int a = -1, b = -1;
        for (int c = 0; c < C && a < 0; c++) if (atA[(size_t)u * C + c] < 0) a = c;
        for (int c = 0; c < C && b < 0; c++) if (atB[(size_t)v * C + c] < 0) b = c;
        if (a < 0 || b < 0) return false;                    // a group with more than 32 edges
        if (atB[(size_t)v * C + a] >= 0) {
            // a is taken at v: flip the a/b alternating path that starts at v (it cannot reach u)
            std::vector<int> path;
            int node = v, side = 1, want = a;
            for (;;) {
                int pe = side ? atB[(size_t)node * C + want] : atA[(size_t)node * C + want];
                if (pe < 0) break;
                path.push_back(pe);
                node = side ? ga[pe] : gb[pe];
                side ^= 1;
                want = (want == a) ? b : a;
            }
            for (int pe : path) {                            // clear, then re-enter with swapped colours
                atA[(size_t)ga[pe] * C + colour[pe]] = -1; atB[(size_t)gb[pe] * C + colour[pe]] = -1;
            }
            for (int pe : path) {
                colour[pe] = (colour[pe] == a) ? b : a;
                atA[(size_t)ga[pe] * C + colour[pe]] = pe; atB[(size_t)gb[pe] * C + colour[pe]] = pe;
            }
        }
        colour[e] = a;
        atA[(size_t)u * C + a] = e; atB[(size_t)v * C + a] = e;
    }
    return true;
}

// Builds the 16-bit records and the shared-memory placement (see the header comment and DeviceGraph::oc_tab).
// Leaves g.oc_tab null (the code then runs on STREAM) when the code is outside the engine's limits.
int onchip_build_tables(ldpc_b200_code *c)
{
    DeviceGraph &g = c->g;
    const int M = c->M, N = c->N, E = c->E, D = g.row_ell_d;
    if (!D || M <= 0 || E <= 0 || N > 65535) return 0;
    g.row_regular = 1;
    for (int m = 0; m < M; m++) if (c->row_ptr[m + 1] - c->row_ptr[m] != D) g.row_regular = 0;
    if (D > 8) g.row_regular = 0;                 // mask-free instances exist for the small row classes only

    // fold position of every edge inside its variable: 0 = last check (the innermost term of Orig.hs:96)
    std::vector<int> fold(E), var_of(E), chk_of(E);
    for (int m = 0; m < M; m++) for (int e = c->row_ptr[m]; e < c->row_ptr[m + 1]; e++) { chk_of[e] = m; var_of[e] = c->col_idx[e]; }
    for (int j = 0; j < N; j++) {
        int q0 = c->col_ptr[j], d = c->col_ptr[j + 1] - q0;
        for (int i = 0; i < d; i++) fold[c->col_edge[q0 + d - 1 - i]] = i;
    }

    // ---- variable positions: a permutation inside every block of 32, banks spread over the check groups ----
    const int nCG = ((M + 31) / 32) * D;                      // check group (m / 32, k)
    std::vector<uint8_t> used((size_t)nCG * 32, 0);
    std::vector<int> pos(N, -1), posvar(N, -1);
    std::vector<std::vector<int>> groups_of(N);
    for (int e = 0; e < E; e++) groups_of[var_of[e]].push_back((chk_of[e] / 32) * D + (e - c->row_ptr[chk_of[e]]));
    for (int h = 0; h * 32 < N; h++) {
        bool taken[32] = {false};
        const int lim = std::min(32, N - h * 32);             // the last block may be short: banks 0 .. lim-1 only
        for (int j = h * 32; j < h * 32 + lim; j++) {
            int best = -1, best_cost = 1 << 30;
            for (int bnk = 0; bnk < lim; bnk++) {
                if (taken[bnk]) continue;
                int cost = 0;
                for (int gq : groups_of[j]) cost += used[(size_t)gq * 32 + bnk];
                if (cost < best_cost) { best_cost = cost; best = bnk; }
            }
            taken[best] = true;
            pos[j] = h * 32 + best; posvar[h * 32 + best] = j;
            for (int gq : groups_of[j]) used[(size_t)gq * 32 + best]++;
        }
    }

    // ---- message slots: 32-colouring of (check group) x (variable group by POSITION, fold position) ----
    const int CF = std::max(c->max_col_deg, 1);
    const int nVG = ((N + 31) / 32) * CF;
    std::vector<int> ga(E), gb(E), colour;
    for (int e = 0; e < E; e++) {
        ga[e] = (chk_of[e] / 32) * D + (e - c->row_ptr[chk_of[e]]);
        gb[e] = (pos[var_of[e]] / 32) * CF + fold[e];
    }
    if (!colour_edges_32(nCG, nVG, ga, gb, colour)) return 0;
    std::vector<int> rank(32, 0), slot(E);
    for (int e = 0; e < E; e++) slot[e] = colour[e] + 32 * rank[colour[e]]++;
    const int SE = 32 * *std::max_element(rank.begin(), rank.end());

    // ---- records (16-bit, BYTE offsets for 4-byte elements; float64 doubles them at use) ----
    //   rowA  [REC4][M] x 8 u16 : the check's record {off(lam) of its D variables, off(ne) of its D messages}, cut
    //                            into 16-byte pieces stored piece-major (threads of consecutive checks read consecutive
    //                            16-byte words: conflict-free 128-bit loads)
    //   rowDeg[M]               : degree (irregular codes only)
    //   colA  [N] x CDR u16     : per variable POSITION {degree, off(ne) of its first CDR-1 messages in fold order}
    //   varPos[N]               : position of variable j (frame in / decisions out)
    if ((size_t)N * 4 > 65535 || (size_t)SE * 4 > 65535) return 0;
    const int REC4 = (2 * D + 7) / 8;
    const int CDR = c->max_col_deg <= 3 ? 4 : 8;
    const int CDs = std::min(c->max_col_deg, CDR - 1);            // staged messages per variable
    auto pad8 = [](std::vector<uint16_t> &v) { while (v.size() % 8) v.push_back(0); };
    std::vector<uint16_t> tab;
    for (int piece = 0; piece < REC4; piece++)
        for (int m = 0; m < M; m++) {
            int b = c->row_ptr[m], d = c->row_ptr[m + 1] - b;
            for (int w = 0; w < 8; w++) {
                int idx = piece * 8 + w;                           // entry of the record: [0, D) positions, [D, 2D) slots
                uint16_t v = 0;
                if (idx < 2 * D && d > 0) {
                    int k = idx % D, e = b + (k < d ? k : d - 1);
                    v = (uint16_t)(4 * (idx < D ? pos[var_of[e]] : slot[e]));
                }
                tab.push_back(v);
            }
        }
    if (!g.row_regular) { for (int m = 0; m < M; m++) tab.push_back((uint16_t)(c->row_ptr[m + 1] - c->row_ptr[m])); pad8(tab); }
    for (int q = 0; q < N; q++) {
        int j = posvar[q], q0 = c->col_ptr[j], d = c->col_ptr[j + 1] - q0;
        tab.push_back((uint16_t)d);
        for (int i = 0; i < CDR - 1; i++) tab.push_back((uint16_t)((d > 0 && i < CDs) ? 4 * slot[c->col_edge[q0 + d - 1 - (i < d ? i : d - 1)]] : 0));
    }
    pad8(tab);
    for (int j = 0; j < N; j++) tab.push_back((uint16_t)pos[j]);
    pad8(tab);
    const int CD = CDs;
    std::vector<int32_t> cs(E);
    for (int q = 0; q < E; q++) cs[q] = slot[c->col_edge[q]];

    LDPC_CUDA(cudaMalloc(&g.oc_tab, sizeof(uint16_t) * tab.size()));
    LDPC_CUDA(cudaMemcpy(g.oc_tab, tab.data(), sizeof(uint16_t) * tab.size(), cudaMemcpyHostToDevice));
    LDPC_CUDA(cudaMalloc(&g.oc_posvar, sizeof(int32_t) * N));
    LDPC_CUDA(cudaMemcpy(g.oc_posvar, posvar.data(), sizeof(int32_t) * N, cudaMemcpyHostToDevice));
    LDPC_CUDA(cudaMalloc(&g.col_slot, sizeof(int32_t) * E));
    LDPC_CUDA(cudaMemcpy(g.col_slot, cs.data(), sizeof(int32_t) * E, cudaMemcpyHostToDevice));
    g.oc_tab_elems = (int)tab.size(); g.oc_se = SE; g.oc_cd = CD; g.oc_cdr = CDR;
    c->oc_slot_of_edge.assign(slot.begin(), slot.end());
    c->oc_pos_of_var.assign(pos.begin(), pos.end());
    return 0;
}

static bool onchip_geometry(const ldpc_b200_code *c, size_t es, OnchipGeom *out)
{
    const DeviceGraph &g = c->g;
    if (!g.oc_tab) return false;
    OnchipGeom geo;
    geo.SE = g.oc_se; geo.CD = g.oc_cd;
    geo.frame_elems = (size_t)2 * c->N + (size_t)geo.SE;
    geo.table_bytes = ((size_t)g.oc_tab_elems * sizeof(uint16_t) + 15) / 16 * 16;
    const size_t frame_bytes = geo.frame_elems * es;
    size_t cta_bytes = ONCHIP_MAX_CTA_BYTES;                                   // two CTAs per SM ...
    if (geo.table_bytes + frame_bytes + 64 > cta_bytes) cta_bytes = 224 * 1024;   // ... or one, for the largest codes
    if (geo.table_bytes + frame_bytes + 64 > cta_bytes) return false;
    int work = std::max(c->M, c->N);
    geo.TPF = std::min(256, std::max(32, ((work + 5) / 6 + 31) / 32 * 32));
    geo.FPC = (int)std::min<size_t>(std::min(8, 512 / geo.TPF), (cta_bytes - 64 - geo.table_bytes) / frame_bytes);
    if (geo.FPC < 1) geo.FPC = 1;
    if (out) *out = geo;
    return true;
}

bool onchip_fits(const ldpc_b200_code *c, int precision) { return onchip_geometry(c, elem_size(precision), nullptr); }

template <typename T, int ALGO, int MAXD, bool REG, int CDR>
static int onchip_launch_t(ldpc_b200_decoder *d, ResidentParams &p, const OnchipGeom &geo, const OnchipLaunch &oc)
{
    auto kern = k_onchip<T, ALGO, MAXD, REG, CDR>;
    const int threads = geo.FPC * geo.TPF;
    const size_t smem = geo.table_bytes + (size_t)geo.FPC * geo.frame_elems * sizeof(T) + geo.FPC * sizeof(unsigned int);
    int blocks_per_sm = 0, n_sm = 0;
    // the dynamic shared-memory limit is a per-device function attribute: raised (never lowered) when a launch needs more;
    // the occupancy of the persistent grid is cached per decoder (one code, one geometry)
    static std::atomic<int> cfg_smem[ONCHIP_MAX_DEVICES];
    const int dev = d->code->device;
    if (dev < 0 || dev >= ONCHIP_MAX_DEVICES) { set_error("device ordinal %d out of range", dev); return LDPC_B200_EINVAL; }
    if (cfg_smem[dev].load(std::memory_order_acquire) < (int)smem) {
        LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        cfg_smem[dev].store((int)smem, std::memory_order_release);
    }
    if (!oc.static_grid) {
        if (!d->oc_blocks_per_sm) {
            LDPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->oc_blocks_per_sm, kern, threads, smem));
            LDPC_CUDA(cudaDeviceGetAttribute(&d->oc_n_sm, cudaDevAttrMultiProcessorCount, dev));
        }
        blocks_per_sm = d->oc_blocks_per_sm; n_sm = d->oc_n_sm;
        if (blocks_per_sm < 1) { d->oc_blocks_per_sm = 0; set_error("ONCHIP kernel does not fit on this device"); return LDPC_B200_EUNSUPPORTED; }
    }
    unsigned int groups = (p.n_frames + geo.FPC - 1) / geo.FPC;
    unsigned int grid;
    if (oc.static_grid) {
        p.work = nullptr;
        grid = groups;
    } else {
        if (!d->res_work) LDPC_CUDA(cudaMalloc(&d->res_work, sizeof(unsigned int)));
        LDPC_CUDA(cudaMemsetAsync(d->res_work, 0, sizeof(unsigned int), oc.stream));
        p.work = d->res_work;
        grid = std::min<unsigned int>(groups, (unsigned int)(blocks_per_sm * n_sm));
    }
    if (grid == 0) return 0;
    kern<<<grid, threads, smem, oc.stream>>>(d->code->g, p, geo);
    d->launches++;
    LDPC_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, int ALGO>
static int onchip_launch_d(ldpc_b200_decoder *d, ResidentParams &p, const OnchipGeom &geo, const OnchipLaunch &oc)
{
    // (regular instances only for the small row classes: a regular code of row degree exactly 18 or 32 takes the masked form)
    const bool reg = d->code->g.row_regular != 0;
    const bool c4 = d->code->g.oc_cdr == 4;
    switch (d->code->g.row_ell_d) {
    case 4:
        if (reg) return c4 ? onchip_launch_t<T, ALGO, 4, true, 4>(d, p, geo, oc) : onchip_launch_t<T, ALGO, 4, true, 8>(d, p, geo, oc);
        return c4 ? onchip_launch_t<T, ALGO, 4, false, 4>(d, p, geo, oc) : onchip_launch_t<T, ALGO, 4, false, 8>(d, p, geo, oc);
    case 8:
        if (reg) return c4 ? onchip_launch_t<T, ALGO, 8, true, 4>(d, p, geo, oc) : onchip_launch_t<T, ALGO, 8, true, 8>(d, p, geo, oc);
        return c4 ? onchip_launch_t<T, ALGO, 8, false, 4>(d, p, geo, oc) : onchip_launch_t<T, ALGO, 8, false, 8>(d, p, geo, oc);
    case 18: return c4 ? onchip_launch_t<T, ALGO, 18, false, 4>(d, p, geo, oc) : onchip_launch_t<T, ALGO, 18, false, 8>(d, p, geo, oc);
    case 32: return c4 ? onchip_launch_t<T, ALGO, 32, false, 4>(d, p, geo, oc) : onchip_launch_t<T, ALGO, 32, false, 8>(d, p, geo, oc);
    default: set_error("ONCHIP engine: check degree above 32"); return LDPC_B200_EUNSUPPORTED;
    }
}

int onchip_launch(ldpc_b200_decoder *d, ResidentParams &p, cudaStream_t stream, bool static_grid)
{
    OnchipGeom geo;
    OnchipLaunch oc = {stream ? stream : d->stream, static_grid};
    if (static_grid) {
        // latency launches: ONE frame per CTA, so that a handful of frames spreads over as many SMs instead of sharing one
        if (!onchip_geometry(d->code, elem_size(d->precision), &geo)) { set_error("ONCHIP engine: the frame state of this code does not fit shared memory"); return LDPC_B200_EUNSUPPORTED; }
        geo.FPC = 1;
    } else
    if (!onchip_geometry(d->code, elem_size(d->precision), &geo)) { set_error("ONCHIP engine: the frame state of this code does not fit shared memory"); return LDPC_B200_EUNSUPPORTED; }
    if (d->algo == LDPC_B200_ALGO_SPA) {
        if (d->precision != LDPC_B200_PREC_FP32) { set_error("ONCHIP engine: sum-product is float32 only"); return LDPC_B200_EUNSUPPORTED; }
        return onchip_launch_d<float, LDPC_B200_ALGO_SPA>(d, p, geo, oc);
    }
    if (d->algo != LDPC_B200_ALGO_MINSUM) { set_error("ONCHIP engine: unsupported algorithm"); return LDPC_B200_EUNSUPPORTED; }
    return d->precision == LDPC_B200_PREC_FP64 ? onchip_launch_d<double, LDPC_B200_ALGO_MINSUM>(d, p, geo, oc)
                                               : onchip_launch_d<float, LDPC_B200_ALGO_MINSUM>(d, p, geo, oc);
}

}  // namespace ldpc
