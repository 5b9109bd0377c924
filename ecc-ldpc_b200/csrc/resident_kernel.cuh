// resident_kernel.cuh -- device code of the RESIDENT engine (see resident_engine.cu for the design notes); included inside
// namespace ldpc by the translation units that instantiate it: resident_engine.cu (production instances),
// resident_latency.cu (one-frame-per-CTA-pair latency instances) and resident_trace.cu (message-tap instances).
#pragma once

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
    static __host__ __device__ constexpr int local_row(int r)       // index of row r inside its part
    {
        int h = C::half_of_row[r];
        for (int q = 0; q < C::HR; q++) if (C::half_rows[h][q] == r) return q;
        return -1;
    }
};

// Variable-phase plan.  RES_VN_PLAN names one of the short-chain plans of qc_tables.inc (VN2, VN3, VN4: S stages;
// VNC1, VNC2, VNC3: singletons stored in the check phase); undefined: the original schedule, one link per stage
// (col_stage, S = maximum column degree).  A plan says for every step q of column c's chain (col_row_desc order:
// innermost link first) in which stage it runs and what it is:
//   chain (0)     load the running value (negL for q = 0), apply the own message, store it back -- as always;
//   singleton (1) store the own message as it is into a slot of its own behind negL (no load);
//   final (2)     load the running value (or negL) AND the column's singleton slots, apply them in chain order, then the
//                 own message: the finished posterior.
// The operands, their order and every rounding are those of the reference's right fold (Orig.hs:95-98) under every plan --
// only who executes a link changes -- so all plans produce the same bits (float64: the reference restatement's); shared-memory operations
// are d loads + d stores per column under every plan; what changes is the number of barriers per iteration, 1 + S.
template <class C> struct VnLegacy {
    static constexpr int S = C::MAXCD, NSLOT = 0;
    static __host__ __device__ constexpr int stage_of(int c, int q)
    {
        for (int t = 0; t < C::MAXCD; t++) if (C::col_stage[c][t] == q) return t;
        return -9;
    }
    static __host__ __device__ constexpr int kind_of(int c, int q) { return q < C::col_deg[c] - 1 ? 0 : 2; }
    static __host__ __device__ constexpr int slot_of(int, int) { return -1; }
};
template <class C, class V> struct VnTable {
    static constexpr int S = V::S, NSLOT = V::NSLOT;
    static __host__ __device__ constexpr int stage_of(int c, int q) { return V::stage[c][q]; }
    static __host__ __device__ constexpr int kind_of(int c, int q) { return V::kind[c][q]; }
    static __host__ __device__ constexpr int slot_of(int c, int q) { return V::slot[c][q]; }
};
// A code struct with a plan attached: the kernels are instantiated on WithPlan<code tables, plan tables> (or on the bare
// code struct: original schedule).
struct VnNone {};
template <class C0, class V> struct WithPlan : C0 { typedef V VN; typedef C0 Base; };
template <class C, class = void> struct VnPlanSel { typedef VnLegacy<C> type; };
template <class C> struct VnPlanSel<C, std::void_t<typename C::VN>> { typedef VnTable<C, typename C::VN> type; };
template <class C0> struct VnPlanSel<WithPlan<C0, VnNone>, void> { typedef VnLegacy<C0> type; };
template <class C> using VnPlan = typename VnPlanSel<C>::type;
// Which plan serves which instance.  RES_VN_PLAN (default VNC2, the measured best: profiles/r02_vn_plans.md) for the
// throughput and message-tap instances -- float32 min-sum, float32 sum-product, float64 --, each overridable
// (RES_VN_PLAN_MIN / _SPA / _F64); the four-part latency instances (one frame per SM: shared memory is no constraint) take
// the one-stage plan VNC1 (RES_VN_PLAN_LAT; one-frame call through the mailbox, 20 iterations: min-sum 60 -> 57 us,
// sum-product 121 -> 110 us).  RES_VN_LEGACY: original schedule everywhere.
#ifndef RES_VN_PLAN
#define RES_VN_PLAN VNC2
#endif
#ifndef RES_VN_LEGACY
#ifndef RES_VN_PLAN_MIN
#define RES_VN_PLAN_MIN RES_VN_PLAN
#endif
#ifndef RES_VN_PLAN_SPA
#define RES_VN_PLAN_SPA RES_VN_PLAN
#endif
#ifndef RES_VN_PLAN_F64
#define RES_VN_PLAN_F64 RES_VN_PLAN
#endif
#ifndef RES_VN_PLAN_LAT
#define RES_VN_PLAN_LAT VNC1
#endif
#endif
// LAT: instance of resident_latency.cu
template <class C0, typename T, int ALGO, bool LAT = false> struct PlanFor { typedef WithPlan<C0, VnNone> type; };
#ifdef RES_VN_PLAN_MIN
template <class C0> struct PlanFor<C0, float, LDPC_B200_ALGO_MINSUM, false> { typedef WithPlan<C0, typename C0::RES_VN_PLAN_MIN> type; };
#endif
#ifdef RES_VN_PLAN_SPA
template <class C0> struct PlanFor<C0, float, LDPC_B200_ALGO_SPA, false> { typedef WithPlan<C0, typename C0::RES_VN_PLAN_SPA> type; };
#endif
#ifdef RES_VN_PLAN_F64
template <class C0, int ALGO> struct PlanFor<C0, double, ALGO, false> { typedef WithPlan<C0, typename C0::RES_VN_PLAN_F64> type; };
#endif
#ifdef RES_VN_PLAN_LAT
template <class C0, int ALGO> struct PlanFor<C0, float, ALGO, true> { typedef WithPlan<C0, typename C0::RES_VN_PLAN_LAT> type; };
#endif
template <class C> struct VnEdge {
    // step of edge (r, k) in its column's chain
    static __host__ __device__ constexpr int step(int r, int k)
    {
        int c = C::row_col[r][k];
        for (int q = 0; q < C::col_deg[c]; q++) if (C::col_row_desc[c][q] == r) return q;
        return -1;
    }
    // the executor's steps of column c (kind 2) all belong to one thread; the LAST of them carries the executor's code
    static __host__ __device__ constexpr int exec_primary(int c)
    {
        int p = -1;
        for (int q = 0; q < C::col_deg[c]; q++) if (VnPlan<C>::kind_of(c, q) == 2) p = q;
        return p;
    }
    static __host__ __device__ constexpr int nchain(int c)
    {
        int n = 0;
        for (int q = 0; q < C::col_deg[c]; q++) n += VnPlan<C>::kind_of(c, q) == 0;
        return n;
    }
    // edge (r, k) belongs to the thread that FINISHES the new posterior of its variable (the executor of the column's plan;
    // originally the last link = the first block row of the column): the same thread reads it for edge (r, k) in the next
    // check phase, which is what the last-link fold uses
    static __host__ __device__ constexpr bool is_final_edge(int r, int k) { return VnPlan<C>::kind_of(C::row_col[r][k], step(r, k)) == 2; }
    // singleton stored in the check phase?
    static __host__ __device__ constexpr bool cn_single(int r, int k)
    {
        int c = C::row_col[r][k], q = step(r, k);
        return VnPlan<C>::kind_of(c, q) == 1 && VnPlan<C>::stage_of(c, q) < 0;
    }
    static __host__ __device__ constexpr int slot(int r, int k) { return VnPlan<C>::slot_of(C::row_col[r][k], step(r, k)); }
};

// Shared address of variable (c = 0, (i+s) mod SZ) given P = address of variable (0, i): the block column is
// BLK = SZ*sizeof(T) bytes and BLK-aligned, so only the low log2(BLK) bits rotate: one add + one bit-select.
// RES_ADDR_IMAD: form the add as an integer multiply-add by an opaque 1 (a __constant__ the compiler cannot fold), so that it
// issues to the FMA pipe (IMAD) instead of the half-rate ALU pipe (VIADD) that the FMNMX trees and the bit-selects live on
// (46 % of the executed instructions are ALU-pipe instructions; `math_pipe_throttle` is 11 % of the stall samples).
// Measured (round 2): jpl.1024.4.5 min-sum 21.0 -> 21.7 Gbit/s with all 20 iterations and 40.5 -> 44.6 at 4.0 dB; jpl.4096.4.5
// min-sum 23.03 -> 23.06 (nothing), sum-product 7.81 -> 7.76.  1: the codes with 32-entry circulants (four frames per CTA), 2: all.
#ifndef RES_ADDR_IMAD
#define RES_ADDR_IMAD 1
#endif
static __constant__ uint32_t k_opaque_one = 1;
template <class C, typename T, int S>
__device__ __forceinline__ uint32_t rot(uint32_t P)
{
    constexpr uint32_t MASK = C::SZ * sizeof(T) - 1;
    if constexpr (S == 0) return P;
    else {
        uint32_t y, z;
        if constexpr (RES_ADDR_IMAD == 2 || (RES_ADDR_IMAD == 1 && C::SZ == 32)) asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(P), "r"(k_opaque_one), "n"(S * (int)sizeof(T)));
        else y = P + S * (uint32_t)sizeof(T);
        asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(z) : "r"(P), "r"(y), "n"(MASK));   // (P & ~MASK) | (y & MASK)
        return z;
    }
}

template <typename T> __device__ __forceinline__ T minus34() { return (T)-0.75; }

// Duplicated posterior columns (RES_DUP): every block column of mu is stored TWICE back to back (2*SZ entries, entry x
// holding variable x mod SZ), so the rotated read of variable (i + s) mod SZ by the thread of check position i is the plain
// address [P + c*2*BLK + s*sizeof(T)] -- an immediate, no add and no bit-select -- for all 99 non-zero-shift circulants in
// the check phase (and the syndrome pass).  The price: the LAST link of a column's chain, which finishes the posterior,
// stores it to both halves (one extra STS per column and iteration; intermediate partial sums live in the first half
// only), and 22 KB more shared memory per float32 frame of jpl.4096.4.5 (90 KB per CTA: still two CTAs per SM).
// MEASURED AND REJECTED (round 2, B200, jpl.4096.4.5 min-sum, 20 iterations): static instructions 4768 -> 4280, but
// 21.6 Gbit/s with it against 22.4 without (with the fused normalisation: 22.6 against 23.1) -- the kernel is bound by
// shared-memory wavefronts as much as by issue slots (2.7 accesses per edge and iteration: 66 % of the LSU roof), and the
// 44 extra stores per iteration cost more than the 200 saved integer instructions.  Kept as a build switch.
#ifndef RES_DUP
#define RES_DUP 0
#endif
template <class C, typename T> struct Lay {
    static constexpr int BLK = C::SZ * (int)sizeof(T);            // one block column, bytes
    static constexpr int CB = RES_DUP ? 2 * BLK : BLK;            // column stride of mu, bytes
    static constexpr int N = C::BC * C::SZ;
    static constexpr int MUE = RES_DUP ? 2 * N : N;               // elements of mu per frame
    static constexpr int NEGL = MUE * (int)sizeof(T);             // byte offset of negL behind mu (block columns of BLK bytes)
    static constexpr int SLOT0 = NEGL + N * (int)sizeof(T);       // byte offset of the singleton slots behind negL (BLK bytes each)
    static constexpr int FS = MUE + N + VnPlan<C>::NSLOT * C::SZ; // elements of mu + negL + singleton slots per frame
    // Does the frame state PLUS a staging row for the next frame's LLRs fit MINB CTAs of FPC frames per SM (227 KB, 1 KB
    // reserved per CTA, 128 bytes of bookkeeping)?  If not -- plans with many singleton slots -- the CTA has no staging row:
    // the next frame's row is only warmed in L2 (cp.async.bulk.prefetch.L2) and read with plain loads when its turn comes.
    static __host__ __device__ constexpr bool staged(int FPC, int MINB)
    {
        return (long long)MINB * ((long long)FPC * (FS + N) * (long long)sizeof(T) + 1024 + 128) <= 227 * 1024;
    }
    static __host__ __device__ constexpr int mu_index(int j) { return RES_DUP ? j + (j & ~(C::SZ - 1)) : j; }   // variable j in mu
    // mu of variable (c, (i + S) mod SZ) for the thread whose P = address of mu entry (0, i)
    template <int c, int S> static __device__ __forceinline__ T lds_mu(uint32_t P)
    {
        if constexpr (RES_DUP) return FB<T>::template lds<c * CB + S * (int)sizeof(T)>(P);
        else return FB<T>::template lds<c * BLK>(rot<C, T, S>(P));
    }
};

// Fused normalisation (float32 min-sum): the message registers hold the UNSCALED leave-one-out reduction r, the reference's
// message being ne' = -3/4 r (Min.hs:79); its two consumers take the factor into a fused multiply-add --
//     x   = ne' + mu   = fma(-3/4, r, mu)        (check phase, Min.hs:80)
//     mu' = acc - ne'  = fma(+3/4, r, acc)       (variable phase, Orig.hs:95-98)
// -- so the multiply costs no instruction of its own (one of 12.3 per edge and iteration) and the product is not rounded
// separately (one rounding instead of two: closer to the float64 reference, no longer bit-identical to "multiply, then
// add").  The float64 instances keep the reference's separate multiply: they reproduce its arithmetic bit for bit.
#ifndef RES_FUSE_SCALE
#define RES_FUSE_SCALE 1
#endif
// Base-2 units (float32 sum-product): the whole frame state -- channel LLRs, posteriors, messages -- is kept in units of ln 2
// (x2 = x * log2 e, scaled once when the frame comes in).  The check update then needs no scaling on either side: the input
// complement is 1 / (1/2 + 2^|x2| / 2) (EX2, FFMA, RCP; before: FFMA, EX2, FADD, RCP), the output magnitude
// lg2((2 - c) / c) is already in the stored unit (before: times ln 2), and the reference's clamp is exactly 54.  Two
// FMA-pipe instructions fewer per edge and iteration; decisions depend on signs only; the message tap scales back.
#ifndef RES_SPA_LOG2
#define RES_SPA_LOG2 1
#endif
#ifndef RES_SPA_STAGGER_NS
#define RES_SPA_STAGGER_NS 0
#endif
#ifndef RES_SPA_REORDER
#define RES_SPA_REORDER 1
#endif
#ifndef RES_MIN_STAGGER_NS
#define RES_MIN_STAGGER_NS 0
#endif
__device__ __forceinline__ float spa_compl2(float x2) { return mufu_rcp(fmaf(mufu_ex2(fabsf(x2)), 0.5f, 0.5f)); }
__device__ __forceinline__ float spa_mag2(float c) { return fminf(mufu_lg2(2.0f - c) - mufu_lg2(c), SPA_CAP_LG2); }
template <typename T, int ALGO> struct Units {
    static constexpr bool LOG2 = RES_SPA_LOG2 && sizeof(T) == 4 && ALGO == LDPC_B200_ALGO_SPA;
    static __device__ __forceinline__ T in(T llr) { if constexpr (LOG2) return __fmul_rn(llr, 1.4426950408889634f); else return llr; }
    static __device__ __forceinline__ T out(T v) { if constexpr (LOG2) return __fmul_rn(v, 0.69314718055994531f); else return v; }
    static __device__ __forceinline__ float compl_of(float x) { if constexpr (LOG2) return spa_compl2(x); else return spa_compl(x); }
    static __device__ __forceinline__ float mag_of(float c) { if constexpr (LOG2) return spa_mag2(c); else return spa_mag(c); }
};

template <typename T, int ALGO> struct MsgForm {
    static constexpr bool FUSED = RES_FUSE_SCALE && sizeof(T) == 4 && ALGO == LDPC_B200_ALGO_MINSUM;
    // x of the next check phase from the stored message s and the (negated) posterior
    static __device__ __forceinline__ T x_of(T s, T mu)
    {
        if constexpr (FUSED) return __fmaf_rn(-0.75f, s, mu); else return FB<T>::add(s, mu);
    }
    // one link of the posterior chain: acc - ne'
    static __device__ __forceinline__ T link(T acc, T s)
    {
        if constexpr (FUSED) return __fmaf_rn(0.75f, s, acc); else return FB<T>::sub(acc, s);
    }
    static __device__ __forceinline__ T store(T others)            // what the check phase leaves in the register
    {
        if constexpr (FUSED) return others; else return FB<T>::mul(minus34<T>(), others);
    }
    static __device__ __forceinline__ T value(T s)                 // the reference's ne' (message tap)
    {
        if constexpr (FUSED) return __fmul_rn(-0.75f, s); else return Units<T, ALGO>::out(s);
    }
};

// A singleton whose plan puts its store into the check phase: right after the row's new messages exist.
template <class C, typename T, int r, int k>
__device__ __forceinline__ void cn_store_single(T v, uint32_t P)
{
    if constexpr (VnEdge<C>::cn_single(r, k))
        FB<T>::template sts<Lay<C, T>::SLOT0 + VnEdge<C>::slot(r, k) * Lay<C, T>::BLK>(rot<C, T, C::row_shift[r][k]>(P), v);
}

// Shared memory holds mu = -lam (the NEGATED posterior) and negL = 0 - L.  Negation is exact and symmetric under
// round-to-nearest, so mu is bit for bit the negation of the reference's lam, with one useful difference: an exact
// zero (punctured LLR, or a sum that cancels) is +0.0 in mu, and mu is never -0.0 (x + y = -0 needs both operands
// -0, and negL = 0 - L is never -0).  Hence hard(lam) = (lam > 0) = sign bit of mu, and the syndrome parity of a
// check row is the xor of the sign bits of its mu's: half a LOP3 per edge instead of a compare per edge.

// Check-node phase for the rows of part H. P = shared address of mu[i]. Returns true if any of this thread's
// checks is unsatisfied by hard(lam).
// FOLD (min-sum instances): for the edges that are the last link of their column's chain, the variable phase (or the
// frame prologue) has already left x = ne + mu in the message register and the sign of mu in fpar[row], so the check
// phase neither loads nor adds for them: 56 of the 156 shared-memory loads of an iteration disappear (44 columns, 12 of them
// with two or more edges in their executor's hands).
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
            if constexpr (!(FOLD && VnEdge<C>::is_final_edge(r, k))) {
                T m = Lay<C, T>::template lds_mu<c, s>(P);
                par ^= F::bits(m);
                ne[base + k] = MsgForm<T, ALGO>::x_of(ne[base + k], m);     // x = -(lam_j - ne[m,j]) = ne + mu   (Min.hs:80)
            }
        });
        unsat |= par;
        if constexpr (ALGO == LDPC_B200_ALGO_MINSUM) {
            // ne'[m,n] = (-3/4) * foldr1 min' [ x_j | j /= n ]   (Min.hs:79-87)
            BoxTree<T, 0, d> tree;
            tree.build([&](auto ki) { return ne[base + decltype(ki)::value]; });
            tree.down_root([&](auto ki, T others) {
                constexpr int k = decltype(ki)::value;
                ne[base + k] = MsgForm<T, ALGO>::store(others);
                cn_store_single<C, T, r, k>(ne[base + k], P);
            });
        } else if constexpr (sizeof(T) == 4) {
            // ne'[m,n] = -2 atanh'(product [ tanh(x_j / 2) | j /= n ])   (Orig.hs:86-91), complement-domain form
            uint32_t sgn = 0;
            static_for<d>([&](auto ki) { sgn ^= __float_as_uint(ne[base + decltype(ki)::value]); });
            sgn = ~sgn;                                                // leading minus times the product of ALL signs
            BoxTree<float, 0, d, OpCompl> tree;
            tree.build([&](auto ki) { return Units<T, ALGO>::compl_of(ne[base + decltype(ki)::value]); });
            tree.down_root([&](auto ki, float others) {
                constexpr int k = decltype(ki)::value;
                // own sign removed from the product: sgn ^ sign(x_k); magnitude >= 0 so its sign bit is free
                ne[base + k] = __uint_as_float(__float_as_uint(Units<T, ALGO>::mag_of(others)) | ((sgn ^ __float_as_uint(ne[base + k])) & 0x80000000u));
                cn_store_single<C, T, r, k>(ne[base + k], P);
            });
        }
    });
    return unsat & F::SIGN;        // non-zero sign bit(s): an unsatisfied check (one bit per packed frame)
}

// RES_VOTE_FIRST (min-sum): the check phase split at the vote -- gather (loads, x = ne + mu, syndrome parity), barrier-vote,
// and only then the leave-one-out trees.  A frame whose syndrome turns out to be satisfied stops before the trees of an
// iteration it does not need (with the fused phase it has always computed one set of messages it never uses); the number
// of barriers is the same, because the trees touch registers only.
// MEASURED AND REJECTED (round 2): jpl.4096.4.5 22.0 against 23.1 Gbit/s at 2.0 dB and 51.6 against 52.2 at 4.0 dB (7.6
// iterations per frame), jpl.1024.4.5 19.8 against 21.1 -- separating the loads from the trees takes away the overlap of the
// LSU and ALU pipes inside a warp, which costs more than the one skipped tree pass per frame saves.  Kept as a build switch.
#ifndef RES_VOTE_FIRST
#define RES_VOTE_FIRST 0
#endif
template <class C, typename T, int H, int ALGO, bool FOLD>
__device__ __forceinline__ typename FB<T>::B cn_gather(T (&ne)[QcMap<C>::NEH], typename FB<T>::B (&fpar)[C::HR], uint32_t P)
{
    typedef FB<T> F;
    typename F::B unsat = 0;
    static_for<C::HR>([&](auto qi) {
        constexpr int r = C::half_rows[H][decltype(qi)::value];
        constexpr int base = QcMap<C>::reg_base(r);
        typename F::B par = 0;
        if constexpr (FOLD) { par = fpar[decltype(qi)::value]; fpar[decltype(qi)::value] = 0; }
        static_for<C::row_deg[r]>([&](auto ki) {
            constexpr int k = decltype(ki)::value;
            if constexpr (!(FOLD && VnEdge<C>::is_final_edge(r, k))) {
                T m = Lay<C, T>::template lds_mu<C::row_col[r][k], C::row_shift[r][k]>(P);
                par ^= F::bits(m);
                ne[base + k] = MsgForm<T, ALGO>::x_of(ne[base + k], m);
            }
        });
        unsat |= par;
    });
    return unsat & F::SIGN;
}
template <class C, typename T, int H, int ALGO>
__device__ __forceinline__ void cn_trees(T (&ne)[QcMap<C>::NEH], uint32_t P)
{
    static_for<C::HR>([&](auto qi) {
        constexpr int r = C::half_rows[H][decltype(qi)::value];
        constexpr int d = C::row_deg[r];
        constexpr int base = QcMap<C>::reg_base(r);
        BoxTree<T, 0, d> tree;                                      // ne'[m,n] = (-3/4) * foldr1 min' [ x_j | j /= n ]   (Min.hs:79-87)
        tree.build([&](auto ki) { return ne[base + decltype(ki)::value]; });
        tree.down_root([&](auto ki, T others) {
            constexpr int k = decltype(ki)::value;
            ne[base + k] = MsgForm<T, ALGO>::store(others);
            cn_store_single<C, T, r, k>(ne[base + k], P);
        });
    });
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
#if RES_SPA_STAGGER_NS > 0
    // experiment: the two warps of a frame that share a scheduler leave the barrier in lockstep and hit the XU pipe with
    // their MUFU bursts at the same time; hold one of them back by a fraction of a row
    if ((((threadIdx.x % (C::NP * C::SZ)) >> 5) & 4) != 0) __nanosleep(RES_SPA_STAGGER_NS);
#endif
    // RES_SPA_REORDER: the two warps of a frame that share a scheduler (warps w and w + 4) leave the barrier in lockstep and
    // would hit the XU pipe with their MUFU bursts at the same time, leaving it idle while both are in their trees.  The
    // upper four warps therefore take the small rows LAST instead of first: a third of a big row's work as a phase offset,
    // at no cost (measured with a sleep first: 0 / 100 / 250 / 500 ns of delay: 8.29 / 8.31 / 8.35 / 8.54 Gbit/s).
    const bool late = RES_SPA_REORDER && ((((threadIdx.x % (C::NP * C::SZ)) >> 5) & 4) != 0);
    auto small_rows = [&]() {
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
                    float m = Lay<C, float>::template lds_mu<C::row_col[r][k], C::row_shift[r][k]>(P);
                    par ^= F::bits(m);
                    ne[base + k] = F::add(ne[base + k], m);
                    sgn ^= F::bits(ne[base + k]);
                });
                unsat |= par;
                sgn = ~sgn;
                BoxTree<float, 0, d, OpCompl> tree;
                tree.build([&](auto ki) { return Units<float, LDPC_B200_ALGO_SPA>::compl_of(ne[base + decltype(ki)::value]); });
                tree.down_root([&](auto ki, float others) {
                    constexpr int k = decltype(ki)::value;
                    ne[base + k] = __uint_as_float(__float_as_uint(Units<float, LDPC_B200_ALGO_SPA>::mag_of(others)) | ((sgn ^ __float_as_uint(ne[base + k])) & 0x80000000u));
                    cn_store_single<C, float, r, k>(ne[base + k], P);
                });
            }
        });
    });
    };
    if (!late) small_rows();
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
                float m = Lay<C, float>::template lds_mu<C::row_col[r][k], C::row_shift[r][k]>(P);
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
            xs[k] = __uint_as_float(__float_as_uint(Units<float, LDPC_B200_ALGO_SPA>::compl_of(xs[k])) | (__float_as_uint(xs[k]) & 0x80000000u));
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
            constexpr int rb = BigRows<C>::row(idx / NB, idx % NB);
            constexpr int base = QcMap<C>::reg_base(rb);
            static_for<D>([&](auto ki) {
                constexpr int k = decltype(ki)::value;
                if constexpr (RES_SPA_LOG2) ne[base + k] = xs[k];            // already in the stored unit
                else ne[base + k] = xs[k] * 0.69314718055994531f;
                cn_store_single<C, float, rb, k>(ne[base + k], P);
            });
        });
    }
    if (late) small_rows();
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
            par ^= F::bits(Lay<C, T>::template lds_mu<C::row_col[r][k], C::row_shift[r][k]>(P));
        });
        unsat |= par;
    });
    return unsat & F::SIGN;
}

// Variable-node stage Q for part H.  Column c builds mu' = (-ne'[m1]) + ((-ne'[m2]) + ... ((-ne'[mk]) + negL)), the
// exact negation of Orig.hs:95-98, from the innermost link (last block row) outwards; VnPlan says which step of which
// column runs in which stage and how (chain link / singleton store / final fold, see VnPlan).  negL sits Lay::NEGL bytes
// after mu, the singleton slots Lay::SLOT0 bytes after it.
// RES_VN_CHUNKS: the columns of a stage are processed in this many groups, each "all loads, then all stores" (inside a
// stage no address is touched by two threads, so the groups need no barrier); bounds the live temporaries of a long stage.
#ifndef RES_VN_CHUNKS
#define RES_VN_CHUNKS 2
#endif
template <class C, typename T, int H, int Q, bool FOLD, int ALGO>
__device__ __forceinline__ void vn_stage(T (&ne)[QcMap<C>::NEH], typename FB<T>::B (&fpar)[C::HR], uint32_t P)
{
    typedef FB<T> F;
    typedef VnPlan<C> V;
    constexpr int BLK = Lay<C, T>::BLK, CB = Lay<C, T>::CB, NEGL = Lay<C, T>::NEGL, SLOT0 = Lay<C, T>::SLOT0;
    constexpr int MAXS = C::MAXCD;                               // indexed by chain step
    constexpr int PER = (C::BC + RES_VN_CHUNKS - 1) / RES_VN_CHUNKS;
    static_for<RES_VN_CHUNKS>([&](auto gi) {
        constexpr int C0 = decltype(gi)::value * PER, C1 = (C0 + PER < C::BC) ? C0 + PER : C::BC;
        // Two passes: every load of the group is issued before the first store.  The accesses are volatile asm and keep
        // their program order, so a fused load-subtract-store per link would serialise the stage on shared-memory latency.
        T acc[PER];
        T sg[PER][MAXS];
        static_for<C1 - C0>([&](auto ci) {
            constexpr int c = C0 + decltype(ci)::value;
            static_for<C::col_deg[c]>([&](auto qi) {
                constexpr int q = decltype(qi)::value;
                constexpr int r = C::col_row_desc[c][q];
                constexpr int kind = V::kind_of(c, q);
                if constexpr (V::stage_of(c, q) == Q && C::half_of_row[r] == H && (kind == 0 || (kind == 2 && q == VnEdge<C>::exec_primary(c)))) {
                    const uint32_t a = rot<C, T, C::row_shift[r][C::col_k_desc[c][q]]>(P);
                    constexpr bool from_negL = (kind == 0) ? (q == 0) : (VnEdge<C>::nchain(c) == 0);
                    acc[c - C0] = from_negL ? F::template lds<NEGL + c * BLK>(a) : F::template lds<c * CB>(a);
                    if constexpr (kind == 2) {
                        static_for<C::col_deg[c]>([&](auto q2i) {
                            constexpr int q2 = decltype(q2i)::value;
                            if constexpr (V::kind_of(c, q2) == 1) sg[c - C0][q2] = F::template lds<SLOT0 + V::slot_of(c, q2) * BLK>(a);
                        });
                    }
                }
            });
        });
        static_for<C1 - C0>([&](auto ci) {
            constexpr int c = C0 + decltype(ci)::value;
            static_for<C::col_deg[c]>([&](auto qi) {
                constexpr int q = decltype(qi)::value;
                constexpr int r = C::col_row_desc[c][q];
                constexpr int kind = V::kind_of(c, q);
                if constexpr (V::stage_of(c, q) == Q && C::half_of_row[r] == H) {
                    constexpr int k = C::col_k_desc[c][q];
                    constexpr int e = QcMap<C>::reg_base(r) + k;
                    const uint32_t a = rot<C, T, C::row_shift[r][k]>(P);
                    if constexpr (kind == 0) {
                        F::template sts<c * CB>(a, MsgForm<T, ALGO>::link(acc[c - C0], ne[e]));
                    } else if constexpr (kind == 1) {
                        F::template sts<SLOT0 + V::slot_of(c, q) * BLK>(a, ne[e]);           // the message as it is; the final link applies it
                    } else if constexpr (q == VnEdge<C>::exec_primary(c)) {
                        // the executor: the rest of the chain in chain order -- slots of the other threads' edges, registers
                        // for its own (edges of this variable in other block rows of this part with the same shift)
                        T val = acc[c - C0];
                        static_for<C::col_deg[c]>([&](auto q2i) {
                            constexpr int q2 = decltype(q2i)::value;
                            if constexpr (V::kind_of(c, q2) == 1) val = MsgForm<T, ALGO>::link(val, sg[c - C0][q2]);
                            else if constexpr (V::kind_of(c, q2) == 2) {
                                constexpr int e2 = QcMap<C>::reg_base(C::col_row_desc[c][q2]) + C::col_k_desc[c][q2];
                                val = MsgForm<T, ALGO>::link(val, ne[e2]);
                            }
                        });
                        F::template sts<c * CB>(a, val);
                        if constexpr (RES_DUP) F::template sts<c * CB + BLK>(a, val);      // the finished posterior: both halves
                        if constexpr (FOLD) {
                            // val is the new mu of this variable, and the executor's edges are this thread's own edges into it
                            static_for<C::col_deg[c]>([&](auto q2i) {
                                constexpr int q2 = decltype(q2i)::value;
                                if constexpr (V::kind_of(c, q2) == 2) {
                                    constexpr int r2 = C::col_row_desc[c][q2];
                                    constexpr int e2 = QcMap<C>::reg_base(r2) + C::col_k_desc[c][q2];
                                    ne[e2] = MsgForm<T, ALGO>::x_of(ne[e2], val);            // x of the next check phase
                                    fpar[QcMap<C>::local_row(r2)] ^= F::bits(val);           // its hard decision, for the next syndrome
                                }
                            });
                        }
                    }
                }
            });
        });
    });
}

// Reference edge index of edge (block row r, check position i, k-th circulant of the row): rows in ascending order,
// ascending column inside a row (the order Orig.hs:82-92 visits the ones of H) -- used by the message tap.
template <class C> struct RefEdge {
    static __host__ __device__ constexpr int row_base(int r)
    {
        int b = 0;
        for (int q = 0; q < r; q++) b += C::row_deg[q] * C::SZ;
        return b;
    }
};

// Decode of one frame by the threads of part H (the iteration loop is instantiated per part so that no dispatch
// happens inside it; all parts execute the same sequence of named barriers).
// H >= 0: instance for part H.  H = -1: one instance for all parts with the part h as a run-time value (the sum-product
// form, whose big-row arithmetic is shared: cn_phase_spa_looped); table-dependent phases branch on h (warp-uniform).
// TRACE: message tap of launch frame 0 (ldpc_b200_decode_trace).  (The output index of the frame -- f or its mailbox slot --
// and the staging row of the packed decisions are recomputed where the frame finishes: nothing extra stays live in the loop.)
template <class C, typename T, int H, int ALGO, bool TRACE, int FPC, bool STAGED>
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
                if constexpr (VnEdge<C>::is_final_edge(r, k)) {
                    T m = Lay<C, T>::template lds_mu<C::row_col[r][k], C::row_shift[r][k]>(P);
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
            if constexpr (H >= 0 && RES_VOTE_FIRST && ALGO == LDPC_B200_ALGO_MINSUM) {
                unsat = ((n == p.max_iters) ? synd_phase<C, T, H>(P) : cn_gather<C, T, H, ALGO, FOLD>(ne, fpar, P)) != 0;
            } else if constexpr (H >= 0) {
#if RES_MIN_STAGGER_NS > 0
                if ((wif & 4) != 0) __nanosleep(RES_MIN_STAGGER_NS);      // experiment: see RES_SPA_REORDER
#endif
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
            const unsigned int fi = p.use_map ? p.slot_map[f] : f;
            // the staging row of this slot (see k_resident): stage - negL = (FPC - slot) * FS - MUE + slot * N; without a staging
            // row: the frame's singleton slots, which are dead once the frame has finished
            static_assert(STAGED || VnPlan<C>::NSLOT * C::SZ * sizeof(T) >= (N / 32) * sizeof(uint32_t), "no room to stage the packed decisions");
            uint32_t *obuf = STAGED ? reinterpret_cast<uint32_t *>(const_cast<T *>(negL) + (size_t)(FPC - (bar - 1)) * Lay<C, T>::FS - Lay<C, T>::MUE + (size_t)(bar - 1) * N)
                                    : reinterpret_cast<uint32_t *>(const_cast<T *>(negL) + N);
            // Warp wif packs the WW consecutive words [wif * WW, (wif + 1) * WW): one ballot per word, lane q keeps word q, so the
            // words leave with ONE coalesced store per warp and the error count is one popc per lane (this loop used to cost a
            // frame ~600 instructions per warp -- three quarters of an iteration; profiles/r02_vn_plans.md).
            constexpr int WW = W / WPF;
            static_assert(W % WPF == 0 && WW <= 32, "decision words must split evenly over the warps of a frame");
            unsigned int errs = 0, myword = 0;
#pragma unroll
            for (int q = 0; q < WW; q++) {
                const int j = 32 * (wif * WW + q) + lane;
                const unsigned int word = __ballot_sync(0xffffffffu, (conv ? mu[Lay<C, T>::mu_index(j)] : negL[j]) < (T)0);   // lam > 0 <=> -lam < 0
                if (lane == q) myword = word;
            }
            if (lane < WW) {
                const int w = wif * WW + lane;
                if (p.done) obuf[w] = myword;
                else if (p.bits) p.bits[(size_t)fi * W + w] = myword;
                if (p.counters) {
                    const int lo = 32 * w;
                    if (lo + 32 <= p.k_info) errs = __popc(myword);
                    else if (lo < p.k_info) errs = __popc(myword & ((1u << (p.k_info - lo)) - 1u));
                }
            }
            if (p.counters) {
                errs = __reduce_add_sync(0xffffffffu, errs);            // lane 0 carries the warp's count (summed up in k_resident)
                if (lane != 0) errs = 0;
            }
            if (p.done) {
                // mailbox: the packed decisions leave in full lines (the arrays are in mapped host memory: one PCIe write per
                // 128 bytes instead of one per word), then the completion flag behind a system-scope fence
                bar_sync(bar, TPF);
                for (int w = tf; w < W; w += TPF) p.bits[(size_t)fi * W + w] = obuf[w];
                if (tf == 0) { if (p.iters) p.iters[fi] = n; if (p.conv) p.conv[fi] = conv ? 1 : 0; }
                __threadfence_system();
                bar_sync(bar, TPF);
                if (tf == 0) *reinterpret_cast<volatile unsigned int *>(p.done + fi) = p.done_seq;
            } else if (tf == 0) {
                if (p.iters) p.iters[fi] = n;
                if (p.conv) p.conv[fi] = conv ? 1 : 0;
            }
            if (p.counters) {
                bool anyerr = bar_or(bar, TPF, errs != 0);
                cnt.berr += errs;
                if (tf == 0) { cnt.frames++; cnt.ferr += anyerr; cnt.iters += n; cnt.unconv += !conv; }
            } else {
                // the decision reads of mu / negL above are done before anybody writes them again: the next stage-0 writes (no
                // early exit) or the next frame's values (a warp packs words that other warps' threads will overwrite)
                bar_sync(bar, TPF);
            }
        }
        if (n == p.max_iters || (done && p.early_exit)) break;
        if constexpr (H >= 0 && RES_VOTE_FIRST && ALGO == LDPC_B200_ALGO_MINSUM) cn_trees<C, T, H, ALGO>(ne, P);

        if constexpr (TRACE) {
            // message tap: ne' of this iteration in reference edge order (Orig.hs:81-92 / Min.hs:75-87), before the last-link
            // fold of the variable phase turns some registers into the next x
            if (f == 0 && p.trace_c2v) {
                T *out = reinterpret_cast<T *>(p.trace_c2v) + (size_t)n * p.trace_ne;
                const int i = tf % C::SZ;
                dispatch_index<0, C::NP>(h, [&](auto hi) {
                    constexpr int HH = decltype(hi)::value;
                    static_for<C::HR>([&](auto qi) {
                        constexpr int r = C::half_rows[HH][decltype(qi)::value];
                        constexpr int d = C::row_deg[r];
                        static_for<d>([&](auto ki) { out[RefEdge<C>::row_base(r) + i * d + decltype(ki)::value] = MsgForm<T, ALGO>::value(ne[Mp::reg_base(r) + decltype(ki)::value]); });
                    });
                });
            }
        }

        // ---- variable phase ----
        static_for<VnPlan<C>::S>([&](auto qi) {
            if constexpr (H >= 0) vn_stage<C, T, H, decltype(qi)::value, FOLD, ALGO>(ne, fpar, P);
            else dispatch_index<0, C::NP>(h, [&](auto hi) { vn_stage<C, T, decltype(hi)::value, decltype(qi)::value, false, ALGO>(ne, fpar, P); });
            bar_sync(bar, TPF);
        });

        if constexpr (TRACE) {
            if (f == 0 && p.trace_lam) {        // lam' = -mu (Orig.hs:95-98); every stage ended behind a barrier
                T *out = reinterpret_cast<T *>(p.trace_lam) + (size_t)n * N;
                for (int j = tf; j < N; j += TPF) out[j] = Units<T, ALGO>::out(-mu[Lay<C, T>::mu_index(j)]);
            }
        }
    }
}

// MINB: resident CTAs per SM the register allocation is bounded for (production: 2 for float32; latency instances: 1,
// which lets the NP = 4 form keep its messages in registers at 512 threads).
template <class C, typename T, int FPC, int ALGO, int MINB, bool TRACE>
__global__ void __launch_bounds__(FPC * C::NP * C::SZ, MINB) k_resident(ResidentParams p)
{
    typedef QcMap<C> Mp;
    typedef FB<T> F;
    constexpr int N = Mp::N, SZ = C::SZ, TPF = C::NP * SZ;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int slot = threadIdx.x / TPF, tf = threadIdx.x % TPF;
    const int h = tf / SZ, i = tf % SZ, lane = threadIdx.x & 31;
    const int bar = slot + 1;
    constexpr int FS = Lay<C, T>::FS, MUE = Lay<C, T>::MUE;
    T *mu = reinterpret_cast<T *>(smem_raw) + (size_t)slot * FS;         // negated posterior, see above (columns duplicated: Lay)
    T *negL = mu + MUE;                                                  // negated channel LLRs
    // staging row of the NEXT frame's LLRs (host-supplied frames only): fetched by one bulk copy while this frame decodes;
    // in mailbox mode (no prefetch) the same row stages the packed decisions
    constexpr bool STAGED = Lay<C, T>::staged(FPC, MINB);               // see Lay::staged
    T *stage = reinterpret_cast<T *>(smem_raw) + (size_t)FPC * FS + (size_t)slot * N;
    unsigned int *s_next = reinterpret_cast<unsigned int *>(reinterpret_cast<T *>(smem_raw) + (size_t)FPC * (FS + (STAGED ? N : 0)));
    uint64_t *mbar = reinterpret_cast<uint64_t *>(s_next + ((FPC + 1) & ~1)) + slot;
    const uint32_t P = (uint32_t)__cvta_generic_to_shared(mu + i);
    auto put_mu = [&](int j, T v) {                       // initial posterior of variable j (both halves of a duplicated column)
        mu[Lay<C, T>::mu_index(j)] = v;
        if constexpr (RES_DUP) mu[Lay<C, T>::mu_index(j) + SZ] = v;
    };
    const bool prefetch = p.llr != nullptr && p.llr_bulk && p.work != nullptr;
    constexpr uint32_t ROWB = N * sizeof(T);

    FrameCounters cnt;
    uint32_t phase = 0;
    // fetch of the row of frame nxt: into the staging row (completes on mbar), or -- no staging row -- into L2 only
    auto fetch_row = [&](unsigned int nxt) {
        if constexpr (STAGED) { mbar_expect_tx(mbar, ROWB); bulk_g2s(stage, reinterpret_cast<const T *>(p.llr) + (size_t)nxt * N, ROWB, mbar); }
        else bulk_prefetch_l2(reinterpret_cast<const T *>(p.llr) + (size_t)nxt * N, ROWB);
    };
    if (prefetch) {
        if (tf == 0) {
            if constexpr (STAGED) mbar_init(mbar, 1);
            const unsigned int nxt = atomicAdd(p.work, 1u);
            s_next[slot] = nxt;
            if (nxt < p.n_frames) fetch_row(nxt);
        }
        bar_sync(bar, TPF);
    }
    for (int round = 0;; round++) {
        unsigned int f;
        if (!p.work) {
            // static assignment (small latency-bound launches): one frame per slot, no work counter to reset or contend on
            if (round) break;
            f = blockIdx.x * FPC + slot;
        } else {
            if (!prefetch) {
                if (tf == 0) s_next[slot] = atomicAdd(p.work, 1u);
                bar_sync(bar, TPF);
            }
            f = s_next[slot];
        }
        if (f >= p.n_frames) break;
        const unsigned int fi = p.use_map ? p.slot_map[f] : f;

        // ---- frame in: negL = 0 - L, mu = negL ----
        if (prefetch) {
            const T *row = stage;
            if constexpr (STAGED) { mbar_wait(mbar, phase); phase ^= 1u; }
            else row = reinterpret_cast<const T *>(p.llr) + (size_t)f * N;     // warmed in L2 while the previous frame decoded
            if constexpr (RES_DUP) {
                for (int j = tf; j < N; j += TPF) { T v = F::sub((T)0, Units<T, ALGO>::in(row[j])); negL[j] = v; put_mu(j, v); }   // 0 - L: never -0.0
            } else {
                // 16 bytes per access: the rows are 16-byte aligned (llr_bulk) and N is a multiple of 32
                constexpr int VE = 16 / (int)sizeof(T);
                struct alignas(16) V16 { T v[VE]; };
                for (int j = tf * VE; j < N; j += TPF * VE) {
                    V16 x = *reinterpret_cast<const V16 *>(row + j);
#pragma unroll
                    for (int q = 0; q < VE; q++) x.v[q] = F::sub((T)0, Units<T, ALGO>::in(x.v[q]));      // 0 - L: never -0.0
                    *reinterpret_cast<V16 *>(negL + j) = x;
                    *reinterpret_cast<V16 *>(mu + j) = x;
                }
            }
            bar_sync(bar, TPF);                 // the staging row has been read by everybody: it can take the next frame
            if (tf == 0) {
                const unsigned int nxt = atomicAdd(p.work, 1u);
                s_next[slot] = nxt;             // read at the top of the next round, many barriers from here
                if (nxt < p.n_frames) fetch_row(nxt);
            }
        } else if (p.llr) {
            const T *src = reinterpret_cast<const T *>(p.llr) + (size_t)fi * N;
            for (int j = tf; j < N; j += TPF) { T v = F::sub((T)0, Units<T, ALGO>::in(src[j])); negL[j] = v; put_mu(j, v); }   // 0 - L: never -0.0
            bar_sync(bar, TPF);
        } else if (p.llr64) {
            // the reference's float64 vector, rounded to T exactly as the host-side conversion would; unpuncture (Utils.hs:55)
            const double *src = p.llr64 + (size_t)fi * p.llr64_stride;
            for (int j = tf; j < N; j += TPF) { T v = F::sub((T)0, j < p.llr64_take ? Units<T, ALGO>::in((T)src[j]) : (T)0); negL[j] = v; put_mu(j, v); }
            bar_sync(bar, TPF);
        } else {
            for (int g = tf; g < N / 4; g += TPF) {
                float nz[4];
                awgn4(p.seed, p.first_frame + f, (uint32_t)g, nz);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    int j = 4 * g + q;
                    T v = F::sub((T)0, (j < p.n_tx) ? Units<T, ALGO>::in((T)awgn_llr(nz[q], 0, p.sigma, p.scale)) : (T)0);
                    negL[j] = v; put_mu(j, v);
                }
            }
            bar_sync(bar, TPF);
        }

        if constexpr (ALGO == LDPC_B200_ALGO_SPA && sizeof(T) == 4 && RES_SPA_LOOPED) {
            decode_frame<C, T, -1, ALGO, TRACE, FPC, STAGED>(p, f, mu, negL, P, bar, tf, lane, cnt, h);
        } else if constexpr (C::NP == 2) {
            if (h == 0) decode_frame<C, T, 0, ALGO, TRACE, FPC, STAGED>(p, f, mu, negL, P, bar, tf, lane, cnt);
            else decode_frame<C, T, 1, ALGO, TRACE, FPC, STAGED>(p, f, mu, negL, P, bar, tf, lane, cnt);
        } else {
            if (h == 0) decode_frame<C, T, 0, ALGO, TRACE, FPC, STAGED>(p, f, mu, negL, P, bar, tf, lane, cnt);
            else if (h == 1) decode_frame<C, T, 1, ALGO, TRACE, FPC, STAGED>(p, f, mu, negL, P, bar, tf, lane, cnt);
            else if (h == 2) decode_frame<C, T, 2, ALGO, TRACE, FPC, STAGED>(p, f, mu, negL, P, bar, tf, lane, cnt);
            else decode_frame<C, T, 3, ALGO, TRACE, FPC, STAGED>(p, f, mu, negL, P, bar, tf, lane, cnt);
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

static constexpr int MAX_DEVICES = 64;

// Launch of one instance.  Function attributes (dynamic shared-memory limit, carve-out) are PER DEVICE, so the "already
// configured" cache is keyed by the device of the code, not by the calling thread; concurrent first launches write the
// same values.  static_grid: one CTA per group of FPC frames and no work counter (small latency-bound launches).
template <class C, typename T, int FPC, int ALGO, int MINB, bool TRACE>
static int launch_instance(ldpc_b200_decoder *d, ResidentParams &p, cudaStream_t stream, bool static_grid)
{
    typedef QcMap<C> Mp;
    auto kern = k_resident<C, T, FPC, ALGO, MINB, TRACE>;
    size_t smem = (size_t)FPC * (Lay<C, T>::FS + (Lay<C, T>::staged(FPC, MINB) ? Mp::N : 0)) * sizeof(T) + ((FPC + 1) & ~1) * sizeof(unsigned int) + FPC * sizeof(uint64_t);
    static_assert((C::SZ * sizeof(T)) <= 1024, "block columns must fit the 1024-byte alignment of the smem base");
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
    unsigned int groups = (p.n_frames + FPC - 1) / FPC;
    unsigned int grid;
    if (static_grid) {
        p.work = nullptr;
        grid = groups;
    } else {
        if (!d->res_work) LDPC_CUDA(cudaMalloc(&d->res_work, sizeof(unsigned int)));
        LDPC_CUDA(cudaMemsetAsync(d->res_work, 0, sizeof(unsigned int), stream));
        p.work = d->res_work;
        grid = std::min<unsigned int>(groups, (unsigned int)(blocks_per_sm * n_sm));
    }
    if (grid == 0) return 0;
    kern<<<grid, FPC * C::NP * C::SZ, smem, stream>>>(p);
    d->launches++;
    LDPC_CUDA(cudaGetLastError());
    return 0;
}
