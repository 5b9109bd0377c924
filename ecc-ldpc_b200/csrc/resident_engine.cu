// resident_engine.cu -- RESIDENT engine: flooding min-sum for the quasi-cyclic JPL AR4JA codes with the
// whole frame state on chip for all iterations.
//
// Why: the reference's state model moves 3E+3N words per frame and iteration; streamed from
// HBM that caps a B200 at ~1.06 M frames/s for jpl.4096.4.5 at 20 iterations.  Here a frame costs HBM only
// its LLRs in and its packed decisions out; the binding resources are instruction issue, the ALU pipe and
// shared-memory bandwidth.
//
// Mapping (SZ = circulant size, NP = threads per check position, one frame = NP*SZ threads):
//   * thread (part, i): check position i of the block rows of its part (tables in qc_tables.inc; the split gives
//     every part the same number of circulants).  Its check->variable messages ne[] live in REGISTERS for the
//     whole decode (one register per edge, compile-time indexed after full unrolling).
//   * the negated posterior mu = -lam and negL = -L live in SHARED memory.  Circulant (r,k) with shift s connects
//     check (r,i) to variable (c, (i+s) mod SZ): a warp reads/writes 32 consecutive words -> conflict-free.
//     Every block column is a SZ*sizeof(T)-aligned block, so the rotated address is formed with one add and one
//     bit-select (LOP3); shifts are compile-time constants and the 57 shift-0 circulants cost nothing.
//   * check phase (Min.hs:75-89): x = ne + mu gathered from smem; syndrome parity = xor of the sign bits of mu
//     (Orig.hs:73-78); leave-one-out reduction with the reference's min' operator (one FMNMX.XORSIGN each) in a
//     binary tree; ne' = -3/4 * (reduction of the others), in place.
//   * vote: barrier.red.or over the frame's threads = "syndrome all zero?"  (Orig.hs:68-69).
//   * variable phase (Orig.hs:95-98): the right fold ne'[m1] + (ne'[m2] + ... (ne'[mk] + L_j)) is built in place in
//     smem, one link per stage from the innermost outwards.  Within a stage every address is touched by exactly
//     one thread; a barrier separates stages; which link of which column runs in which stage is scheduled offline
//     (tools/gen_qc_tables.py) so that the parts carry equal work in every stage.  Same association as the
//     reference, so the fp64 instance reproduces the reference bit for bit.
//     A stage issues all its loads before its first store (the accesses are ordered volatile asm).
//   * last-link fold (min-sum): the thread that executes the LAST link of a column's chain finishes the new posterior
//     of that variable and owns one of its edges, so it keeps x = ne' + mu and the sign of mu for that edge in
//     registers: 44 of the 156 check-phase loads of an iteration never happen.
//   * early exit per frame; frames are handed out by an atomic work counter (persistent CTAs); host-supplied frames:
//     the NEXT frame's LLR row is fetched by one bulk copy (cp.async.bulk + mbarrier) while the current one decodes.
//   * sum-product (float32) runs the same phases in the complement domain, with one shared copy of the degree-18 row
//     arithmetic (cn_phase_spa_looped) because the fully unrolled body did not fit the instruction cache.
//
// fp32 is the production instance; fp64 exists to prove the index logic bit-exactly.
#include "common.cuh"
#include "philox.cuh"
#include <atomic>
#include <utility>

// RES_PARTS: threads per check position (2 or 4). More parts = fewer message registers per thread = more resident
// warps to hide shared-memory latency, at the price of more barrier participants.
#ifndef RES_PARTS
#define RES_PARTS 2
#endif
// RES_SPA_LOOPED: sum-product check phase with one shared copy of the big-row arithmetic (see cn_phase_spa_looped)
#ifndef RES_SPA_LOOPED
#define RES_SPA_LOOPED 1
#endif
// RES_FOLD: min-sum check phase takes x of the last-link edges from the variable phase (see cn_phase)
#ifndef RES_FOLD
#define RES_FOLD 1
#endif

namespace ldpc {

#include "qc_tables.inc"
#include "bp_tree.cuh"
#include "resident_common.cuh"
#include "tma.cuh"

// Compile-time layout of the per-thread message registers: the rows of a part, in order, each with row_deg slots.
template <class C> struct QcMap {
    static __host__ __device__ constexpr int reg_base(int r)
    {
        int h = C::half_of_row[r], b = 0;
        for (int q = 0; q < C::HR; q++) {
            int rr = C::half_rows[h][q];
            if (rr == r) return b;
            b += C::row_deg[rr];
        }
        return -1;
    }
    static __host__ __device__ constexpr int part_edges(int h)
    {
        int b = 0;
        for (int q = 0; q < C::HR; q++) b += C::row_deg[C::half_rows[h][q]];
        return b;
    }
    static __host__ __device__ constexpr int max_part_edges()
    {
        int m = 0;
        for (int h = 0; h < C::NP; h++) m = part_edges(h) > m ? part_edges(h) : m;
        return m;
    }
    static constexpr int NEH = max_part_edges();
    static constexpr int N = C::BC * C::SZ;
    // edge (r, k) is the LAST link of its column's chain (the first block row of the column): the thread that finishes
    // the new posterior of that variable is also the one that reads it for edge (r, k) in the next check phase
    static __host__ __device__ constexpr bool is_final_edge(int r, int k)
    {
        int c = C::row_col[r][k];
        return C::col_row_desc[c][C::col_deg[c] - 1] == r;
    }
    static __host__ __device__ constexpr int local_row(int r)       // index of row r inside its part
    {
        int h = C::half_of_row[r];
        for (int q = 0; q < C::HR; q++) if (C::half_rows[h][q] == r) return q;
        return -1;
    }
};

// Shared address of variable (c = 0, (i+s) mod SZ) given P = address of variable (0, i): the block column is
// BLK = SZ*sizeof(T) bytes and BLK-aligned, so only the low log2(BLK) bits rotate: one add + one bit-select.
template <class C, typename T, int S>
__device__ __forceinline__ uint32_t rot(uint32_t P)
{
    constexpr uint32_t MASK = C::SZ * sizeof(T) - 1;
    if constexpr (S == 0) return P;
    else {
        uint32_t y = P + S * (uint32_t)sizeof(T), z;
        asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(z) : "r"(P), "r"(y), "n"(MASK));   // (P & ~MASK) | (y & MASK)
        return z;
    }
}

template <typename T> __device__ __forceinline__ T minus34() { return (T)-0.75; }

// Shared memory holds mu = -lam (the NEGATED posterior) and negL = 0 - L.  Negation is exact and symmetric under
// round-to-nearest, so mu is bit for bit the negation of the reference's lam, with one useful difference: an exact
// zero (punctured LLR, or a sum that cancels) is +0.0 in mu, and mu is never -0.0 (x + y = -0 needs both operands
// -0, and negL = 0 - L is never -0).  Hence hard(lam) = (lam > 0) = sign bit of mu, and the syndrome parity of a
// check row is the xor of the sign bits of its mu's: half a LOP3 per edge instead of a compare per edge.

// Check-node phase for the rows of part H. P = shared address of mu[i]. Returns true if any of this thread's
// checks is unsatisfied by hard(lam).
// FOLD (min-sum instances): for the edges that are the last link of their column's chain, the variable phase (or the
// frame prologue) has already left x = ne + mu in the message register and the sign of mu in fpar[row], so the check
// phase neither loads nor adds for them: 44 of the 156 shared-memory loads of an iteration disappear.
template <class C, typename T, int H, int ALGO, bool FOLD>
__device__ __forceinline__ typename FB<T>::B cn_phase(T (&ne)[QcMap<C>::NEH], typename FB<T>::B (&fpar)[C::HR], uint32_t P)
{
    typedef FB<T> F;
    constexpr int BLK = C::SZ * sizeof(T);
    typename F::B unsat = 0;
    static_for<C::HR>([&](auto qi) {
        constexpr int r = C::half_rows[H][decltype(qi)::value];
        constexpr int d = C::row_deg[r];
        constexpr int base = QcMap<C>::reg_base(r);
        typename F::B par = 0;
        if constexpr (FOLD) { par = fpar[decltype(qi)::value]; fpar[decltype(qi)::value] = 0; }
        static_for<d>([&](auto ki) {
            constexpr int k = decltype(ki)::value;
            constexpr int c = C::row_col[r][k];
            constexpr int s = C::row_shift[r][k];
            if constexpr (!(FOLD && QcMap<C>::is_final_edge(r, k))) {
                T m = F::template lds<c * BLK>(rot<C, T, s>(P));
                par ^= F::bits(m);
                ne[base + k] = F::add(ne[base + k], m);     // x = -(lam_j - ne[m,j]) = ne + mu   (Min.hs:80)
            }
        });
        unsat |= par;
        if constexpr (ALGO == LDPC_B200_ALGO_MINSUM) {
            // ne'[m,n] = (-3/4) * foldr1 min' [ x_j | j /= n ]   (Min.hs:79-87)
            BoxTree<T, 0, d> tree;
            tree.build([&](auto ki) { return ne[base + decltype(ki)::value]; });
            tree.down_root([&](auto ki, T others) { ne[base + decltype(ki)::value] = F::mul(minus34<T>(), others); });
        } else if constexpr (sizeof(T) == 4) {
            // ne'[m,n] = -2 atanh'(product [ tanh(x_j / 2) | j /= n ])   (Orig.hs:86-91), complement-domain form
            uint32_t sgn = 0;
            static_for<d>([&](auto ki) { sgn ^= __float_as_uint(ne[base + decltype(ki)::value]); });
            sgn = ~sgn;                                                // leading minus times the product of ALL signs
            BoxTree<float, 0, d, OpCompl> tree;
            tree.build([&](auto ki) { return spa_compl(ne[base + decltype(ki)::value]); });
            tree.down_root([&](auto ki, float others) {
                constexpr int k = decltype(ki)::value;
                // own sign removed from the product: sgn ^ sign(x_k); magnitude >= 0 so its sign bit is free
                ne[base + k] = __uint_as_float(__float_as_uint(spa_mag(others)) | ((sgn ^ __float_as_uint(ne[base + k])) & 0x80000000u));
            });
        }
    });
    return unsat & F::SIGN;        // non-zero sign bit(s): an unsatisfied check (one bit per packed frame)
}

// Exclusive compile-time dispatch on a run-time index: a binary tree of (warp-uniform) branches whose leaves call
// fn(integral_constant<idx>).
template <int LO, int HI, class Fn>
__device__ __forceinline__ void dispatch_index(int sel, Fn &&fn)
{
    if constexpr (HI - LO == 1) fn(std::integral_constant<int, LO>{});
    else {
        constexpr int MID = (LO + HI) / 2;
        if (sel < MID) dispatch_index<LO, MID>(sel, fn);
        else dispatch_index<MID, HI>(sel, fn);
    }
}

// The "big" rows of a code are those of maximum degree (18 for the AR4JA codes); every part carries the same number.
template <class C> struct BigRows {
    static __host__ __device__ constexpr int count(int h)
    {
        int n = 0;
        for (int q = 0; q < C::HR; q++) n += C::row_deg[C::half_rows[h][q]] == C::MAXD;
        return n;
    }
    static __host__ __device__ constexpr int row(int h, int j)
    {
        for (int q = 0; q < C::HR; q++)
            if (C::row_deg[C::half_rows[h][q]] == C::MAXD && j-- == 0) return C::half_rows[h][q];
        return -1;
    }
    static __host__ __device__ constexpr bool uniform()
    {
        for (int h = 1; h < C::NP; h++) if (count(h) != count(0)) return false;
        return true;
    }
    static constexpr int NB = count(0);
};

// Sum-product check phase with ONE copy of the arithmetic of a big row.  Fully unrolled, the sum-product loop body
// is ~100 KB of SASS (30 instructions per edge, 156 circulants): it does not fit the 32 KB instruction cache and the
// kernel was fetch-bound (ncu: 43 % of the stall samples `no_instruction`).  Here the per-row code is only what
// depends on the circulant table -- gather x = ne + mu at the start, the sign/magnitude merge into ne' at the end --
// selected by a warp-uniform branch tree; the complement / tree / logarithm arithmetic of a degree-MAXD row exists once
// and is shared by all big rows of all parts.  The small rows (degree 3) stay inline.
template <class C>
__device__ __forceinline__ uint32_t cn_phase_spa_looped(float (&ne)[QcMap<C>::NEH], uint32_t P, int h)
{
    typedef FB<float> F;
    constexpr int BLK = C::SZ * sizeof(float), D = C::MAXD, NB = BigRows<C>::NB;
    static_assert(BigRows<C>::uniform(), "parts must carry the same number of maximum-degree rows");
    uint32_t unsat = 0;
    dispatch_index<0, C::NP>(h, [&](auto hi) {
        constexpr int H = decltype(hi)::value;
        static_for<C::HR>([&](auto qi) {
            constexpr int r = C::half_rows[H][decltype(qi)::value];
            constexpr int d = C::row_deg[r];
            constexpr int base = QcMap<C>::reg_base(r);
            if constexpr (d != D) {
                uint32_t par = 0, sgn = 0;
                static_for<d>([&](auto ki) {
                    constexpr int k = decltype(ki)::value;
                    float m = F::template lds<C::row_col[r][k] * BLK>(rot<C, float, C::row_shift[r][k]>(P));
                    par ^= F::bits(m);
                    ne[base + k] = F::add(ne[base + k], m);
                    sgn ^= F::bits(ne[base + k]);
                });
                unsat |= par;
                sgn = ~sgn;
                BoxTree<float, 0, d, OpCompl> tree;
                tree.build([&](auto ki) { return spa_compl(ne[base + decltype(ki)::value]); });
                tree.down_root([&](auto ki, float others) {
                    constexpr int k = decltype(ki)::value;
                    ne[base + k] = __uint_as_float(__float_as_uint(spa_mag(others)) | ((sgn ^ __float_as_uint(ne[base + k])) & 0x80000000u));
                });
            }
        });
    });
#pragma unroll 1
    for (int j = 0; j < NB; j++) {
        float xs[D];
        uint32_t par = 0;
        const int sel = h * NB + j;
        dispatch_index<0, C::NP * NB>(sel, [&](auto ci) {
            constexpr int idx = decltype(ci)::value;
            constexpr int r = BigRows<C>::row(idx / NB, idx % NB);
            constexpr int base = QcMap<C>::reg_base(r);
            static_for<D>([&](auto ki) {
                constexpr int k = decltype(ki)::value;
                float m = F::template lds<C::row_col[r][k] * BLK>(rot<C, float, C::row_shift[r][k]>(P));
                par ^= F::bits(m);
                xs[k] = F::add(ne[base + k], m);                  // x = ne + mu   (Orig.hs:84)
            });
        });
        unsat |= par;
        // ne'[m,n] = -2 atanh'(product [ tanh(x_j / 2) | j /= n ])   (Orig.hs:86-91), complement-domain form.
        // The sign of x_k rides in the sign bit of its (non-negative) complement: the tree reads magnitudes through
        // operand modifiers, so x itself is dead after this point (18 registers less at the widest place).
        uint32_t sgn = 0;
        static_for<D>([&](auto ki) { sgn ^= __float_as_uint(xs[decltype(ki)::value]); });
        // cap with the sign "leading minus times the product of ALL signs" (still to be multiplied by sign(x_k))
        const uint32_t cap_s = (~sgn & 0x80000000u) | __float_as_uint(SPA_CAP_LG2);
        static_for<D>([&](auto ki) {
            constexpr int k = decltype(ki)::value;
            xs[k] = __uint_as_float(__float_as_uint(spa_compl(xs[k])) | (__float_as_uint(xs[k]) & 0x80000000u));
        });
        BoxTree<float, 0, D, OpComplAbs> tree;
        tree.build([&](auto ki) { return xs[decltype(ki)::value]; });
        tree.down_root([&](auto ki, float others) {
            constexpr int k = decltype(ki)::value;
            // min(|lg2((2-c)/c)|, cap) with sign = sign(cap_s) xor sign(x_k): one LOP3 + one FMNMX.XORSIGN
            xs[k] = F::box(spa_mag_lg2(others), __uint_as_float(cap_s ^ (__float_as_uint(xs[k]) & 0x80000000u)));
        });
        dispatch_index<0, C::NP * NB>(sel, [&](auto ci) {
            constexpr int idx = decltype(ci)::value;
            constexpr int base = QcMap<C>::reg_base(BigRows<C>::row(idx / NB, idx % NB));
            static_for<D>([&](auto ki) { ne[base + decltype(ki)::value] = xs[decltype(ki)::value] * 0.69314718055994531f; });
        });
    }
    return unsat & F::SIGN;
}

// Syndrome only (used for the test after the last iteration).
template <class C, typename T, int H>
__device__ __forceinline__ typename FB<T>::B synd_phase(uint32_t P)
{
    typedef FB<T> F;
    constexpr int BLK = C::SZ * sizeof(T);
    typename F::B unsat = 0;
    static_for<C::HR>([&](auto qi) {
        constexpr int r = C::half_rows[H][decltype(qi)::value];
        constexpr int d = C::row_deg[r];
        typename F::B par = 0;
        static_for<d>([&](auto ki) {
            constexpr int k = decltype(ki)::value;
            par ^= F::bits(F::template lds<C::row_col[r][k] * BLK>(rot<C, T, C::row_shift[r][k]>(P)));
        });
        unsat |= par;
    });
    return unsat & F::SIGN;
}

// Variable-node stage Q for part H.  Column c builds mu' = (-ne'[m1]) + ((-ne'[m2]) + ... ((-ne'[mk]) + negL)), the
// exact negation of Orig.hs:95-98, one link per stage from the innermost (last block row) outwards; col_stage
// says which link runs in which stage.  negL sits N*sizeof(T) bytes after mu.
template <class C, typename T, int H, int Q, bool FOLD>
__device__ __forceinline__ void vn_stage(T (&ne)[QcMap<C>::NEH], typename FB<T>::B (&fpar)[C::HR], uint32_t P)
{
    typedef FB<T> F;
    constexpr int BLK = C::SZ * sizeof(T);
    constexpr int NEGL = QcMap<C>::N * sizeof(T);
    // Two passes: every load of the stage is issued before the first store.  The accesses are volatile asm and keep
    // their program order, so a fused load-subtract-store per link would serialise the stage on shared-memory latency.
    T acc[C::BC];
    static_for<C::BC>([&](auto ci) {
        constexpr int c = decltype(ci)::value;
        constexpr int step = C::col_stage[c][Q];          // which link of column c's chain runs in this stage
        if constexpr (step >= 0) {
            constexpr int r = C::col_row_desc[c][step];
            if constexpr (C::half_of_row[r] == H) {
                constexpr int k = C::col_k_desc[c][step];
                uint32_t a = rot<C, T, C::row_shift[r][k]>(P);
                acc[c] = (step == 0) ? F::template lds<NEGL + c * BLK>(a) : F::template lds<c * BLK>(a);
            }
        }
    });
    static_for<C::BC>([&](auto ci) {
        constexpr int c = decltype(ci)::value;
        constexpr int step = C::col_stage[c][Q];
        if constexpr (step >= 0) {
            constexpr int r = C::col_row_desc[c][step];
            if constexpr (C::half_of_row[r] == H) {
                constexpr int k = C::col_k_desc[c][step];
                constexpr int e = QcMap<C>::reg_base(r) + k;
                uint32_t a = rot<C, T, C::row_shift[r][k]>(P);
                const T val = F::sub(acc[c], ne[e]);
                F::template sts<c * BLK>(a, val);
                if constexpr (FOLD && step == C::col_deg[c] - 1) {
                    // last link: val is the new mu of this variable, and edge e is this thread's own edge into it
                    ne[e] = F::add(ne[e], val);                                  // x of the next check phase
                    fpar[QcMap<C>::local_row(r)] ^= F::bits(val);                // its hard decision, for the next syndrome
                }
            }
        }
    });
}

// Decode of one frame by the threads of part H (the iteration loop is instantiated per part so that no dispatch
// happens inside it; all parts execute the same sequence of named barriers).
// H >= 0: instance for part H.  H = -1: one instance for all parts with the part h as a run-time value (the sum-product
// form, whose big-row arithmetic is shared: cn_phase_spa_looped); table-dependent phases branch on h (warp-uniform).
template <class C, typename T, int H, int ALGO>
__device__ __forceinline__ void decode_frame(const ResidentParams &p, unsigned int f, const T *mu, const T *negL, uint32_t P,
                                             int bar, int tf, int lane, FrameCounters &cnt, int h = H)
{
    typedef QcMap<C> Mp;
    constexpr int N = Mp::N, TPF = C::NP * C::SZ, WPF = TPF / 32, W = N / 32;
    const int wif = tf >> 5;
    constexpr bool FOLD = (H >= 0) && ALGO == LDPC_B200_ALGO_MINSUM && RES_FOLD;
    T ne[Mp::NEH];
    typename FB<T>::B fpar[C::HR];
#pragma unroll
    for (int e = 0; e < Mp::NEH; e++) ne[e] = (T)0;               // orig_ne = 0 (Orig.hs:64-65)
#pragma unroll
    for (int q = 0; q < C::HR; q++) fpar[q] = 0;
    if constexpr (FOLD) {
        // prologue of the fold: x = 0 + mu and the sign of mu for the last-link edges (what the variable phase leaves
        // behind from the second iteration on)
        static_for<C::HR>([&](auto qi) {
            constexpr int r = C::half_rows[H][decltype(qi)::value];
            static_for<C::row_deg[r]>([&](auto ki) {
                constexpr int k = decltype(ki)::value;
                if constexpr (Mp::is_final_edge(r, k)) {
                    T m = FB<T>::template lds<C::row_col[r][k] * C::SZ * (int)sizeof(T)>(rot<C, T, C::row_shift[r][k]>(P));
                    ne[Mp::reg_base(r) + k] = m;
                    fpar[decltype(qi)::value] ^= FB<T>::bits(m);
                }
            });
        });
    }

    bool done = false;
    for (int n = 0;; n++) {
        bool finished_now = false, conv = false;
        if (!(done && n == p.max_iters)) {
            bool unsat;
            if constexpr (H >= 0) {
                unsat = ((n == p.max_iters) ? synd_phase<C, T, H>(P) : cn_phase<C, T, H, ALGO, FOLD>(ne, fpar, P)) != 0;
            } else {
                typename FB<T>::B u = 0;
                if (n == p.max_iters) dispatch_index<0, C::NP>(h, [&](auto hi) { u = synd_phase<C, T, decltype(hi)::value>(P); });
                else u = cn_phase_spa_looped<C>(ne, P, h);
                unsat = u != 0;
            }
            bool any = bar_or(bar, TPF, unsat);
            if (!done && (!any || n == p.max_iters)) { finished_now = true; conv = !any; }
        }
        if (finished_now) {
            // syndrome ok -> hard(lam); iterations exhausted -> hard(orig_lam)   (Orig.hs:68-71)
            done = true;
            const T *src = conv ? mu : negL;
            unsigned int errs = 0;
            for (int w = wif; w < W; w += WPF) {
                unsigned int word = __ballot_sync(0xffffffffu, src[32 * w + lane] < (T)0);   // lam > 0 <=> -lam < 0
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
                bar_sync(bar, TPF);          // snapshot reads of mu before the next stage-0 writes
            }
        }
        if (n == p.max_iters || (done && p.early_exit)) break;

        // ---- variable phase ----
        static_for<C::MAXCD>([&](auto qi) {
            if constexpr (H >= 0) vn_stage<C, T, H, decltype(qi)::value, FOLD>(ne, fpar, P);
            else dispatch_index<0, C::NP>(h, [&](auto hi) { vn_stage<C, T, decltype(hi)::value, decltype(qi)::value, false>(ne, fpar, P); });
            bar_sync(bar, TPF);
        });
    }
}

template <class C, typename T, int FPC, int ALGO>
__global__ void __launch_bounds__(FPC * C::NP * C::SZ, sizeof(T) == 4 ? 2 : 1) k_resident(ResidentParams p)
{
    typedef QcMap<C> Mp;
    typedef FB<T> F;
    constexpr int N = Mp::N, SZ = C::SZ, TPF = C::NP * SZ;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int slot = threadIdx.x / TPF, tf = threadIdx.x % TPF;
    const int h = tf / SZ, i = tf % SZ, lane = threadIdx.x & 31;
    const int bar = slot + 1;
    T *mu = reinterpret_cast<T *>(smem_raw) + (size_t)slot * 2 * N;      // negated posterior, see above
    T *negL = mu + N;                                                    // negated channel LLRs
    // staging row of the NEXT frame's LLRs (host-supplied frames only): fetched by one bulk copy while this frame decodes
    T *stage = reinterpret_cast<T *>(smem_raw) + (size_t)FPC * 2 * N + (size_t)slot * N;
    unsigned int *s_next = reinterpret_cast<unsigned int *>(reinterpret_cast<T *>(smem_raw) + (size_t)FPC * 3 * N);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(s_next + ((FPC + 1) & ~1)) + slot;
    const uint32_t P = (uint32_t)__cvta_generic_to_shared(mu + i);
    const bool prefetch = p.llr != nullptr && p.llr_bulk;
    constexpr uint32_t ROWB = N * sizeof(T);

    FrameCounters cnt;
    uint32_t phase = 0;
    if (prefetch) {
        if (tf == 0) {
            mbar_init(mbar, 1);
            const unsigned int nxt = atomicAdd(p.work, 1u);
            s_next[slot] = nxt;
            if (nxt < p.n_frames) { mbar_expect_tx(mbar, ROWB); bulk_g2s(stage, reinterpret_cast<const T *>(p.llr) + (size_t)nxt * N, ROWB, mbar); }
        }
        bar_sync(bar, TPF);
    }
    for (;;) {
        if (!prefetch) {
            if (tf == 0) s_next[slot] = atomicAdd(p.work, 1u);
            bar_sync(bar, TPF);
        }
        const unsigned int f = s_next[slot];
        if (f >= p.n_frames) break;

        // ---- frame in: negL = 0 - L, mu = negL ----
        if (prefetch) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
            for (int j = tf; j < N; j += TPF) { T v = F::sub((T)0, stage[j]); negL[j] = v; mu[j] = v; }   // 0 - L: never -0.0
            bar_sync(bar, TPF);                 // the staging row has been read by everybody: it can take the next frame
            if (tf == 0) {
                const unsigned int nxt = atomicAdd(p.work, 1u);
                s_next[slot] = nxt;             // read at the top of the next round, many barriers from here
                if (nxt < p.n_frames) { mbar_expect_tx(mbar, ROWB); bulk_g2s(stage, reinterpret_cast<const T *>(p.llr) + (size_t)nxt * N, ROWB, mbar); }
            }
        } else if (p.llr) {
            const T *src = reinterpret_cast<const T *>(p.llr) + (size_t)f * N;
            for (int j = tf; j < N; j += TPF) { T v = F::sub((T)0, src[j]); negL[j] = v; mu[j] = v; }   // 0 - L: never -0.0
            bar_sync(bar, TPF);
        } else {
            for (int g = tf; g < N / 4; g += TPF) {
                float nz[4];
                awgn4(p.seed, p.first_frame + f, (uint32_t)g, nz);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    int j = 4 * g + q;
                    T v = F::sub((T)0, (j < p.n_tx) ? (T)awgn_llr(nz[q], 0, p.sigma, p.scale) : (T)0);
                    negL[j] = v; mu[j] = v;
                }
            }
            bar_sync(bar, TPF);
        }

        if constexpr (ALGO == LDPC_B200_ALGO_SPA && sizeof(T) == 4 && RES_SPA_LOOPED) {
            decode_frame<C, T, -1, ALGO>(p, f, mu, negL, P, bar, tf, lane, cnt, h);
        } else if constexpr (C::NP == 2) {
            if (h == 0) decode_frame<C, T, 0, ALGO>(p, f, mu, negL, P, bar, tf, lane, cnt);
            else decode_frame<C, T, 1, ALGO>(p, f, mu, negL, P, bar, tf, lane, cnt);
        } else {
            if (h == 0) decode_frame<C, T, 0, ALGO>(p, f, mu, negL, P, bar, tf, lane, cnt);
            else if (h == 1) decode_frame<C, T, 1, ALGO>(p, f, mu, negL, P, bar, tf, lane, cnt);
            else if (h == 2) decode_frame<C, T, 2, ALGO>(p, f, mu, negL, P, bar, tf, lane, cnt);
            else decode_frame<C, T, 3, ALGO>(p, f, mu, negL, P, bar, tf, lane, cnt);
        }
        // all smem reads of this frame are done before the next frame's values arrive: decode_frame ends behind a barrier
    }

    if (p.counters) {
        if (lane == 0 && cnt.berr) atomicAdd(&p.counters[2], cnt.berr);
        if (tf == 0 && cnt.frames) {
            atomicAdd(&p.counters[0], cnt.frames); atomicAdd(&p.counters[1], cnt.ferr);
            atomicAdd(&p.counters[3], cnt.iters); atomicAdd(&p.counters[4], cnt.unconv);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

#if RES_PARTS == 2
typedef QcJpl1024P2 QcJpl1024;
typedef QcJpl4096P2 QcJpl4096;
#else
typedef QcJpl1024P4 QcJpl1024;
typedef QcJpl4096P4 QcJpl4096;
#endif

template <class C>
static bool matches(const ldpc_b200_code *c)
{
    if (c->sz != C::SZ || c->brows != C::BR || c->bcols != C::BC) return false;
    for (int r = 0; r < C::BR; r++) {
        int k = 0;
        for (int b = 0; b < C::BC; b++) {
            int s = c->shifts[r * C::BC + b];
            if (s < 0) continue;
            if (k >= C::row_deg[r] || C::row_col[r][k] != b || C::row_shift[r][k] != s) return false;
            k++;
        }
        if (k != C::row_deg[r]) return false;
    }
    return true;
}

int resident_kind_for(const ldpc_b200_code *c)
{
    if (c->sz == 0) return 0;
    if (matches<QcJpl1024>(c)) return 1;
    if (matches<QcJpl4096>(c)) return 2;
    return 0;
}

// which on-chip kernel serves this decoder: the compiled specialisation of its code, or (no specialisation, or the
// ONCHIP engine asked for explicitly) the run-time table kernel of onchip_engine.cu
int resident_kind_of(const ldpc_b200_decoder *d)
{
    return d->engine == LDPC_B200_ENGINE_ONCHIP ? 3 : d->code->resident_kind;
}

bool resident_supports(const ldpc_b200_decoder *d)
{
    // min-sum: fp32 and fp64; sum-product: fp32 only (fp64 sum-product is the STREAM engine's literal form)
    if (d->algo != LDPC_B200_ALGO_SPA && d->algo != LDPC_B200_ALGO_MINSUM) return false;   // the arraylet-order statements live on STREAM
    if (!(d->algo == LDPC_B200_ALGO_MINSUM || d->precision == LDPC_B200_PREC_FP32)) return false;
    if (resident_kind_of(d) == 3) return onchip_fits(d->code, d->precision);
    return d->code->resident_kind != 0;
}

// float64 rows of `take` values at a stride of `stride` -> float rows of N values, zero beyond `take` (the punctured
// positions, Utils.hs:55)
template <typename T>
__global__ void k_convert_f64(const double *__restrict__ in, T *__restrict__ out, size_t n, int N, int stride, int take)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const size_t f = idx / (size_t)N;
    const int j = (int)(idx - f * (size_t)N);
    out[idx] = j < take ? (T)in[f * (size_t)stride + j] : (T)0;
}

static constexpr int MAX_DEVICES = 64;

template <class C, typename T, int FPC, int ALGO>
static int launch(ldpc_b200_decoder *d, ResidentParams &p)
{
    typedef QcMap<C> Mp;
    auto kern = k_resident<C, T, FPC, ALGO>;
    size_t smem = (size_t)FPC * 3 * Mp::N * sizeof(T) + ((FPC + 1) & ~1) * sizeof(unsigned int) + FPC * sizeof(uint64_t);
    static_assert((C::SZ * sizeof(T)) <= 1024, "block columns must fit the 1024-byte alignment of the smem base");
    // Function attributes (dynamic shared-memory limit, carve-out) are PER DEVICE, so the "already configured" cache is
    // keyed by the device of the code, not by the calling thread; concurrent first launches write the same values.
    static std::atomic<int> cfg_bps[MAX_DEVICES], cfg_sm[MAX_DEVICES];
    const int dev = d->code->device;
    if (dev < 0 || dev >= MAX_DEVICES) { set_error("device ordinal %d out of range", dev); return LDPC_B200_EINVAL; }
    int blocks_per_sm = cfg_bps[dev].load(std::memory_order_acquire), n_sm = cfg_sm[dev].load(std::memory_order_relaxed);
    if (!blocks_per_sm) {
        LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        LDPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, FPC * C::NP * C::SZ, smem));
        LDPC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        if (blocks_per_sm < 1) { set_error("RESIDENT kernel does not fit on this device"); return LDPC_B200_EUNSUPPORTED; }
        cfg_sm[dev].store(n_sm, std::memory_order_relaxed);
        cfg_bps[dev].store(blocks_per_sm, std::memory_order_release);
    }
    if (!d->res_work) LDPC_CUDA(cudaMalloc(&d->res_work, sizeof(unsigned int)));
    LDPC_CUDA(cudaMemsetAsync(d->res_work, 0, sizeof(unsigned int), d->stream));
    p.work = d->res_work;
    unsigned int groups = (p.n_frames + FPC - 1) / FPC;
    unsigned int grid = std::min<unsigned int>(groups, (unsigned int)(blocks_per_sm * n_sm));
    if (grid == 0) return 0;
    kern<<<grid, FPC * C::NP * C::SZ, smem, d->stream>>>(p);
    d->launches++;
    LDPC_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, int ALGO>
static int launch_kind(ldpc_b200_decoder *d, ResidentParams &p)
{
    switch (resident_kind_of(d)) {
    case 1: return launch<QcJpl1024, T, 4, ALGO>(d, p);
    case 2: return launch<QcJpl4096, T, 1, ALGO>(d, p);
    case 3: return onchip_launch(d, p);
    default: set_error("no RESIDENT specialisation for this code"); return LDPC_B200_EUNSUPPORTED;
    }
}

static int launch_prec(ldpc_b200_decoder *d, ResidentParams &p)
{
    if (d->algo == LDPC_B200_ALGO_SPA) return launch_kind<float, LDPC_B200_ALGO_SPA>(d, p);
    return d->precision == LDPC_B200_PREC_FP64 ? launch_kind<double, LDPC_B200_ALGO_MINSUM>(d, p)
                                               : launch_kind<float, LDPC_B200_ALGO_MINSUM>(d, p);
}

int resident_decode(ldpc_b200_decoder *d, int max_iters, int early_exit, int n_frames, const void *d_llr,
                    bool llr_is_double, uint32_t *d_bits_packed, int32_t *d_iters, uint8_t *d_conv, int src_stride, int src_take)
{
    if (!resident_supports(d)) { set_error("RESIDENT engine does not support this decoder"); return LDPC_B200_EUNSUPPORTED; }
    int N = d->code->N;
    const void *src = d_llr;
    if (llr_is_double && d->precision == LDPC_B200_PREC_FP32) {
        size_t need = (size_t)d->max_batch * N * sizeof(float);
        if (d->res_conv_bytes < need) {
            if (d->res_conv_buf) cudaFree(d->res_conv_buf);
            d->res_conv_buf = nullptr; d->res_conv_bytes = 0;
            LDPC_CUDA(cudaMalloc(&d->res_conv_buf, need));
            d->res_conv_bytes = need;
        }
        size_t n = (size_t)n_frames * N;
        const int stride = src_take > 0 ? src_stride : N, take = src_take > 0 ? src_take : N;
        k_convert_f64<float><<<(unsigned)((n + 255) / 256), 256, 0, d->stream>>>((const double *)d_llr, (float *)d->res_conv_buf, n, N, stride, take);
        d->launches++;
        src = d->res_conv_buf;
    } else if (!llr_is_double && d->precision == LDPC_B200_PREC_FP64) {
        set_error("fp64 decoder needs float64 LLRs");
        return LDPC_B200_EINVAL;
    } else if (src_take > 0) {
        set_error("packed source rows need the float64 -> float32 conversion pass");
        return LDPC_B200_EINVAL;
    }
    ResidentParams p = {};
    
    p.llr = src; p.bits = d_bits_packed; p.iters = d_iters; p.conv = d_conv; p.counters = nullptr;
    {
        const size_t row = (size_t)N * elem_size(d->precision);
        p.llr_bulk = ((uintptr_t)src % 16 == 0 && row % 16 == 0) ? 1 : 0;     // cp.async.bulk needs 16-byte aligned rows
    }
    p.n_frames = (unsigned int)n_frames; p.max_iters = max_iters; p.early_exit = early_exit;
    p.k_info = 0; p.n_tx = N;
    return launch_prec(d, p);
}

int resident_simulate(ldpc_b200_decoder *d, int max_iters, int early_exit, float sigma, float llr_scale,
                      int k_info, int n_tx, uint64_t seed, uint64_t first_frame, uint64_t n_frames,
                      uint64_t *d_counters)
{
    if (!resident_supports(d)) { set_error("RESIDENT engine does not support this decoder"); return LDPC_B200_EUNSUPPORTED; }
    const uint64_t chunk = 1ull << 30;
    for (uint64_t f0 = 0; f0 < n_frames; f0 += chunk) {
        ResidentParams p = {};
        
        p.llr = nullptr; p.bits = nullptr; p.iters = nullptr; p.conv = nullptr;
        p.counters = (unsigned long long *)d_counters;
        p.n_frames = (unsigned int)std::min<uint64_t>(chunk, n_frames - f0);
        p.max_iters = max_iters; p.early_exit = early_exit;
        p.sigma = sigma; p.scale = llr_scale; p.n_tx = n_tx; p.k_info = k_info;
        p.seed = seed; p.first_frame = first_frame + f0;
        int rc = launch_prec(d, p);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace ldpc
