// cluster_sync_probe.cu -- how much a barrier costs when one frame is spread over a thread-block cluster instead of one CTA.
// The RESIDENT variable phase needs a barrier per stage (7 per iteration); this measures, on the device, the cycles of
//   (a) bar.sync over one CTA of 512 threads (the NP = 4 latency kernel),
//   (b) barrier.cluster.arrive.release + wait.acquire over clusters of 2 and 4 CTAs of 128 / 256 threads,
// each as the average over 2000 back-to-back barriers with one dependent shared-memory store + load in between.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_sync_probe cluster_sync_probe.cu && ./cluster_sync_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void k_cta(long long *out, int iters)
{
    __shared__ float s[512];
    s[threadIdx.x] = threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    float acc = 0.f;
    for (int i = 0; i < iters; i++) {
        s[threadIdx.x] = acc + 1.0f;
        __syncthreads();
        acc += s[(threadIdx.x + 37) % blockDim.x];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)acc; }
}

__global__ void k_cluster(long long *out, int iters)
{
    cg::cluster_group cl = cg::this_cluster();
    __shared__ float s[512];
    s[threadIdx.x] = threadIdx.x;
    cl.sync();
    const unsigned peer = (cl.block_rank() + 1) % cl.num_blocks();
    float *remote = cl.map_shared_rank(s, peer);
    long long t0 = clock64();
    float acc = 0.f;
    for (int i = 0; i < iters; i++) {
        s[threadIdx.x] = acc + 1.0f;
        cl.sync();                                   // barrier.cluster.arrive.release + barrier.cluster.wait.acquire
        acc += remote[(threadIdx.x + 37) % blockDim.x];   // the partial sum written by another CTA of the cluster (DSMEM)
        cl.sync();                                   // (reads done before the next round's writes, as the stage barrier would)
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && cl.block_rank() == 0) { out[0] = t1 - t0; out[1] = (long long)acc; }
}

int main()
{
    long long *d, h[2];
    cudaMalloc(&d, 16);
    const int iters = 2000;
    for (int threads : {128, 256, 512}) {
        k_cta<<<1, threads>>>(d, iters);
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("CTA of %3d threads: %.0f cycles per (store, bar.sync, load) round\n", threads, (double)h[0] / iters);
    }
    for (int csz : {2, 4}) for (int threads : {128, 256}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(csz); cfg.blockDim = dim3(threads);
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = csz; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_cluster, d, iters);
        if (e != cudaSuccess) { printf("cluster %d x %d: %s\n", csz, threads, cudaGetErrorString(e)); continue; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("cluster of %d CTAs x %3d threads: %.0f cycles per (store, cluster.sync, remote load, cluster.sync) round = %.0f per cluster.sync incl. the DSMEM load\n",
               csz, threads, (double)h[0] / iters, (double)h[0] / iters / 2);
    }
    return 0;
}
