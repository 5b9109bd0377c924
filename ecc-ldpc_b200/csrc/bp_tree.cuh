// bp_tree.cuh -- compile-time building blocks shared by the RESIDENT and STREAM engines (included inside namespace ldpc):
// static_for, exact-rounding arithmetic / bit access per floating type, and the leave-one-out binary tree used by the
// check-node updates (min' of Min.hs:89 and the complement-domain product of the fp32 sum-product).
#pragma once
// (the including file has #include <utility> outside the namespace)

template <int... Is, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F &&f)
{
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    static_for_impl(std::make_integer_sequence<int, N>{}, f);
}

// Arithmetic with explicit rounding (no FMA contraction), bit access, and shared-memory access by 32-bit shared
// address + compile-time byte offset (so that the block-column offset rides in the instruction's immediate).
template <typename T> struct FB;
template <> struct FB<float> {
    typedef uint32_t B;
    static constexpr B SIGN = 0x80000000u;
    static __device__ __forceinline__ B bits(float x) { return __float_as_uint(x); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    // min' of Min.hs:89 in ONE instruction: FMNMX.XORSIGN = min(|a|,|b|) carrying sign(a) xor sign(b)
    static __device__ __forceinline__ float box(float a, float b)
    {
        float r;
        asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
        return r;
    }
    template <int OFF> static __device__ __forceinline__ float lds(uint32_t addr)
    {
        float v;
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
        return v;
    }
    template <int OFF> static __device__ __forceinline__ void sts(uint32_t addr, float v)
    {
        asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "f"(v));
    }
};
template <> struct FB<double> {
    typedef unsigned long long B;
    static constexpr B SIGN = 0x8000000000000000ull;
    static __device__ __forceinline__ B bits(double x) { return (B)__double_as_longlong(x); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double box(double a, double b)
    {
        return __longlong_as_double((long long)(bits(fmin(fabs(a), fabs(b))) | ((bits(a) ^ bits(b)) & SIGN)));
    }
    template <int OFF> static __device__ __forceinline__ double lds(uint32_t addr)
    {
        double v;
        asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
        return v;
    }
    template <int OFF> static __device__ __forceinline__ void sts(uint32_t addr, double v)
    {
        asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "d"(v));
    }
};

// Leave-one-out "box" reduction over the D messages of a check row with a binary tree: `build` reduces subtrees
// bottom-up, `down` hands every leaf the reduction of everything outside it.  The operator is the reference's
// min' (Min.hs:89: sign product times minimum magnitude), exact in floating point and associative, so the result
// equals foldr1 min' over the other entries bit for bit.  ~2.6 ops per edge at depth log2(D).
template <typename T> struct OpBox { static __device__ __forceinline__ T f(T a, T b) { return FB<T>::box(a, b); } };
template <typename T> struct OpAdd { static __device__ __forceinline__ T f(T a, T b) { return FB<T>::add(a, b); } };

template <typename T, int LO, int HI, class Op = OpBox<T>> struct BoxTree {
    static constexpr int MID = (LO + HI) / 2;
    BoxTree<T, LO, MID, Op> l;
    BoxTree<T, MID, HI, Op> r;
    T m;
    template <class A> __device__ __forceinline__ void build(const A &a)
    {
        l.build(a); r.build(a);
        m = Op::f(l.m, r.m);
    }
    template <class Fn> __device__ __forceinline__ void down(T o, Fn &&emit) const
    {
        l.down(Op::f(o, r.m), emit);
        r.down(Op::f(o, l.m), emit);
    }
    template <class Fn> __device__ __forceinline__ void down_root(Fn &&emit) const
    {
        l.down(r.m, emit);
        r.down(l.m, emit);
    }
};
template <typename T, int LO, class Op> struct BoxTree<T, LO, LO + 1, Op> {
    T m;
    template <class A> __device__ __forceinline__ void build(const A &a) { m = a(std::integral_constant<int, LO>{}); }
    template <class Fn> __device__ __forceinline__ void down(T o, Fn &&emit) const { emit(std::integral_constant<int, LO>{}, o); }
};

#include "spa_fp32.cuh"
// 1 - (1 - a)(1 - b) in the form a + b (1 - a): still one FADD + one FFMA and no cancellation, and EXACTLY 1 whenever a or b
// is exactly 1 (a = 1: b * 0 + 1; b = 1: (1 - a) + a, where 1 - a is exact for a >= 1/2 and otherwise off by at most 2^-25,
// which the final rounding at 1 absorbs).  That matters: x = 0 (a punctured or erased position, Utils.hs:55) has complement
// exactly 1, and the reference's product then makes every OTHER message of the check exactly 0 -- so a variable that only
// hears from such checks keeps lam = 0 and decides 0 (`hard` = (> 0)).  With the symmetric form a + b - a b the "1" came out
// as 1 - 6e-8, the messages as +-1e-7 with a definite sign, and such variables decided at random: on random codes with
// erased positions the float32 sum-product agreed with the restatement on only 87-92 % of the frames
// (tests/test_gpu_parity.py::test_random_user_codes_all_engines; the JPL codes hide it because their punctured columns
// hear from checks with a single punctured neighbour in the first iteration).
struct OpCompl { static __device__ __forceinline__ float f(float a, float b) { return fmaf(b, 1.0f - a, a); } };
// the same on magnitudes (|.| is an operand modifier of FADD / FFMA): leaves may carry a payload in their sign bit
struct OpComplAbs { static __device__ __forceinline__ float f(float a, float b) { return fmaf(fabsf(b), 1.0f - fabsf(a), fabsf(a)); } };
// lg2((2 - c) / c) = 2 atanh(1 - c) / ln 2 (uncapped), and the reference's ceiling 2 * 18.714973875118524 in the same unit
__device__ __forceinline__ float spa_mag_lg2(float c) { c = fabsf(c); return mufu_lg2(2.0f - c) - mufu_lg2(c); }
#define SPA_CAP_LG2 54.0f   // 2 * 18.714973875118524 = ln((2 - 2^-53) / 2^-53) = 54 ln 2 to within 1e-16


// One check row: x[0..MAXD) holds x_j = -(lam_j - ne[m,j]) for the row's d edges and the neutral element (+inf) in the
// padding slots k >= d (FULL: d == MAXD, no padding); emit(k, ne'[m,k]) is called for every slot, the caller drops
// k >= d.  Min-sum: ne' = (-3/4) * foldr1 min' [ x_j | j /= n ] (Min.hs:79-89; min' is exact and associative).
// Sum-product (fp32 production form, see spa_fp32.cuh): complements c_j = 1 - tanh(|x_j|/2) (pad: x = +inf -> c = 0, the
// identity) combined leave-one-out in the tree; the sign of x_j rides in the sign bit of its complement; magnitude
// lg2((2-c)/c) capped at the reference's clamp and signed by one FMNMX.XORSIGN; times ln 2.
template <typename T, int ALGO, int MAXD, class Emit>
__device__ __forceinline__ void cn_row_core(const T (&x)[MAXD], Emit &&emit)
{
    if constexpr (ALGO == LDPC_B200_ALGO_SPA) {
        uint32_t sgn = 0;
#pragma unroll
        for (int k = 0; k < MAXD; k++) sgn ^= __float_as_uint((float)x[k]);
        const uint32_t cap_s = (~sgn & 0x80000000u) | __float_as_uint(SPA_CAP_LG2);
        float leaf[MAXD];
#pragma unroll
        for (int k = 0; k < MAXD; k++)
            leaf[k] = __uint_as_float(__float_as_uint(spa_compl((float)x[k])) | (__float_as_uint((float)x[k]) & 0x80000000u));
        BoxTree<float, 0, MAXD, OpComplAbs> tree;
        tree.build([&](auto ki) { return leaf[decltype(ki)::value]; });
        tree.down_root([&](auto ki, float others) {
            constexpr int k = decltype(ki)::value;
            const float w = FB<float>::box(spa_mag_lg2(others), __uint_as_float(cap_s ^ (__float_as_uint(leaf[k]) & 0x80000000u)));
            emit(ki, (T)(w * 0.69314718055994531f));
        });
    } else {
        BoxTree<T, 0, MAXD> tree;
        tree.build([&](auto ki) { return x[decltype(ki)::value]; });
        tree.down_root([&](auto ki, T others) { emit(ki, FB<T>::mul((T)-0.75, others)); });
    }
}
