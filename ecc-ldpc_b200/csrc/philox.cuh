// philox.cuh -- Philox4x32-10 (Salmon et al., SC'11) and the BPSK/AWGN LLR generator built on it.
//
// The reference gets its frames from ecc-manifold's tester (external to the repo: random message ->
// encode -> BPSK + AWGN -> LLR); this is the device-side equivalent used by
// ldpc_b200_generate_llr / ldpc_b200_simulate.  Stream layout: key = seed, counter = (group, 0, frame_lo,
// frame_hi) where `group` indexes 4 consecutive codeword positions, so a frame's noise depends only on
// (seed, global frame index) -- never on batching or on how frames are sharded over GPUs.
#pragma once
#include <stdint.h>

namespace ldpc {

struct Philox4 { uint32_t v[4]; };

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

#ifdef __CUDACC__
// Four N(0,1) samples for positions 4*group .. 4*group+3 of one frame (Box-Muller, fp32).
__device__ __forceinline__ void awgn4(uint64_t seed, uint64_t frame, uint32_t group, float n[4])
{
    Philox4 r = philox4x32_10(group, 0u, (uint32_t)frame, (uint32_t)(frame >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
    const float two_m24 = 5.9604644775390625e-08f;   // 2^-24
#pragma unroll
    for (int p = 0; p < 2; p++) {
        float u1 = ((float)(r.v[2 * p] >> 8) + 0.5f) * two_m24;        // (0,1)
        float u2 = ((float)(r.v[2 * p + 1] >> 8) + 0.5f) * two_m24;
        float rad = sqrtf(-2.0f * logf(u1));
        float s, c;
        sincospif(2.0f * u2, &s, &c);
        n[2 * p] = rad * c;
        n[2 * p + 1] = rad * s;
    }
}

// LLR of one position: y = (2b-1) + sigma*n ; llr = (2/sigma^2) * y   (bit 1 -> +1, so llr > 0 decides 1)
__device__ __forceinline__ float awgn_llr(float noise, int bit, float sigma, float llr_scale)
{
    float y = __fmaf_rn(sigma, noise, bit ? 1.0f : -1.0f);
    return __fmul_rn(llr_scale, y);
}
#endif

}  // namespace ldpc
