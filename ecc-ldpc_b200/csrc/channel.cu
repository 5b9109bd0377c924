// channel.cu -- device-side channel frames and error counting: the part of ecc-manifold's tester
// (external; call sites main/Main.hs:48, Utils.hs:59-75) that surrounds `decode` in a BER run.
#include "common.cuh"
#include "philox.cuh"

namespace ldpc {

// One thread per group of 4 positions of one frame. Positions >= n_tx are punctured: LLR 0 (Utils.hs:55).
template <typename T>
__global__ void k_generate(T *__restrict__ llr, const uint32_t *__restrict__ cw, int N, int n_tx, int F,
                           uint64_t seed, uint64_t first_frame, float sigma, float scale)
{
    int groups = (N + 3) / 4;
    int gidx = blockIdx.y * blockDim.x + threadIdx.x;
    int f = blockIdx.x;
    if (gidx >= groups || f >= F) return;
    float n[4];
    awgn4(seed, first_frame + (uint64_t)f, (uint32_t)gidx, n);
    int W = (N + 31) / 32;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int j = gidx * 4 + q;
        if (j >= N) break;
        float v = 0.0f;
        if (j < n_tx) {
            int bit = cw ? (int)((cw[(size_t)f * W + (j >> 5)] >> (j & 31)) & 1u) : 0;
            v = awgn_llr(n[q], bit, sigma, scale);
        }
        llr[(size_t)f * N + j] = (T)v;
    }
}

int channel_generate(ldpc_b200_decoder *d, float sigma, float llr_scale, int n_tx, uint64_t seed,
                     uint64_t first_frame, int n_frames, const uint32_t *d_codewords, void *d_llr)
{
    int N = d->code->N;
    if (n_frames <= 0) return 0;
    dim3 grid(n_frames, ((N + 3) / 4 + 127) / 128);     // frames on x: y is limited to 65535
    if (d->precision == LDPC_B200_PREC_FP64)
        k_generate<double><<<grid, 128, 0, d->stream>>>((double *)d_llr, d_codewords, N, n_tx, n_frames, seed, first_frame, sigma, llr_scale);
    else
        k_generate<float><<<grid, 128, 0, d->stream>>>((float *)d_llr, d_codewords, N, n_tx, n_frames, seed, first_frame, sigma, llr_scale);
    d->launches++;
    LDPC_CUDA(cudaGetLastError());
    return 0;
}

// counters: [0] frames [1] frame errors [2] bit errors (first k_info bits, all-zero codeword) [3] sum iters [4] unconverged
__global__ void k_count(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ ref, const int32_t *__restrict__ iters,
                        const uint8_t *__restrict__ conv, int F, int W, int k_info,
                        unsigned long long *__restrict__ counters)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long fe = 0, be = 0, it = 0, un = 0, fr = 0;
    if (f < F) {
        int full = k_info >> 5, rem = k_info & 31;
        int e = 0;
        for (int w = 0; w < full; w++) e += __popc(bits[(size_t)f * W + w] ^ (ref ? ref[(size_t)f * W + w] : 0u));
        if (rem) e += __popc((bits[(size_t)f * W + full] ^ (ref ? ref[(size_t)f * W + full] : 0u)) & ((1u << rem) - 1u));
        be = e; fe = e != 0; it = iters[f]; un = conv[f] == 0; fr = 1;
    }
    for (int o = 16; o; o >>= 1) {
        fe += __shfl_down_sync(0xffffffffu, fe, o);
        be += __shfl_down_sync(0xffffffffu, be, o);
        it += __shfl_down_sync(0xffffffffu, it, o);
        un += __shfl_down_sync(0xffffffffu, un, o);
        fr += __shfl_down_sync(0xffffffffu, fr, o);
    }
    if ((threadIdx.x & 31) == 0 && fr) {
        atomicAdd(&counters[0], fr); atomicAdd(&counters[1], fe); atomicAdd(&counters[2], be);
        atomicAdd(&counters[3], it); atomicAdd(&counters[4], un);
    }
}

int channel_count(ldpc_b200_decoder *d, int n_frames, int k_info, const uint32_t *d_bits_packed,
                  const uint32_t *d_ref_packed, const int32_t *d_iters, const uint8_t *d_conv, uint64_t *d_counters)
{
    if (n_frames <= 0) return 0;
    k_count<<<(n_frames + 127) / 128, 128, 0, d->stream>>>(d_bits_packed, d_ref_packed, d_iters, d_conv, n_frames,
                                                           words_per_frame(d->code->N), k_info,
                                                           (unsigned long long *)d_counters);
    d->launches++;
    LDPC_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_unpack(const uint32_t *__restrict__ packed, uint8_t *__restrict__ bytes, int N, int W, size_t total)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    size_t f = i / N;
    int j = (int)(i - f * N);
    bytes[i] = (uint8_t)((packed[f * W + (j >> 5)] >> (j & 31)) & 1u);
}

int unpack_bits(ldpc_b200_decoder *d, int n_frames, const uint32_t *d_bits_packed, uint8_t *d_bytes)
{
    int N = d->code->N;
    size_t total = (size_t)n_frames * N;
    if (!total) return 0;
    k_unpack<<<(unsigned)((total + 255) / 256), 256, 0, d->stream>>>(d_bits_packed, d_bytes, N, words_per_frame(N), total);
    d->launches++;
    LDPC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ldpc
