// spa_fp32.cuh -- fp32 sum-product arithmetic shared by the STREAM and RESIDENT engines (included inside namespace ldpc).
#pragma once
// Sum-product in fp32 (production form of Orig.hs:81-92) in the COMPLEMENT domain.  The reference multiplies
// t_j = tanh(x_j/2); in fp32 that product saturates (tanhf = 1 beyond |x| ~ 18 and the atanh' clamp then jumps to
// 37.43).  We carry c_j = 1 - |t_j| = 2 e^-|x_j| / (1 + e^-|x_j|) instead, which keeps full relative precision however
// confident the message is, and combine with
//     1 - (1 - ca)(1 - cb) = ca + cb - ca cb        (no cancellation: the result is >= max(ca, cb)),
// leave-one-out by the same binary tree as min-sum.  With c = 1 - prod_{j/=n} |t_j| the outgoing magnitude is
//     2 atanh(1 - c) = ln((2 - c) / c),
// capped at the reference's own ceiling 2 * 18.714973875118524 (Utils.hs:113-117); signs are handled as a parity.
// x = 0 gives c = 1 (t = 0) and therefore 0 on all other edges, as the product does.
// Four MUFU per edge (ex2, rcp, 2 x lg2), all .ftz: no denormal fix-up code and no data-dependent slow path.
// The complement is formed as 1 / (1/2 + e^|x|/2): one FFMA, EX2, FADD, RCP and no cancellation anywhere.
__device__ __forceinline__ float mufu_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float spa_compl(float x)            // 1 - tanh(|x|/2) = 2 / (1 + e^|x|) = 1 / (1/2 + e^|x| / 2)
{
    const float half_e = mufu_ex2(fmaf(fabsf(x), 1.4426950408889634f, -1.0f));     // e^|x| / 2; +inf beyond |x| ~ 89 -> c = 0
    return mufu_rcp(0.5f + half_e);
}
__device__ __forceinline__ float spa_mag(float c)              // 2 atanh(1 - c), capped
{
    const float m = (mufu_lg2(2.0f - c) - mufu_lg2(c)) * 0.69314718055994531f;
    return fminf(m, 37.429947750237048f);
}
