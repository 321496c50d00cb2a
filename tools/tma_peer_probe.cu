// Probe: does cp.async.bulk (the 1-D TMA copy of k_st3) read NVLink peer memory?  Single process, 2 devices.
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_peer tools/tma_peer_probe.cu && /tmp/tma_peer   -> 0 mismatches (B200 x2)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const double *src, double *out, int n) {
    extern __shared__ __align__(128) unsigned char sm[];
    double *buf = (double *)sm;
    uint64_t *bar = (uint64_t *)(sm + n * 8);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(n * 8) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf)), "l"(src), "r"(n * 8), "r"(s32(bar)) : "memory");
    }
    asm volatile("{\n.reg .pred P1;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra DN;\nbra LW;\nDN:\n}" ::"r"(s32(bar)) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}
int main() {
    int n = 4096, nd = 0;
    cudaGetDeviceCount(&nd);
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    double *src1, *out0, *h = new double[n];
    cudaSetDevice(1);
    cudaMalloc(&src1, n * 8);
    for (int i = 0; i < n; ++i) h[i] = 1.5 * i + 7;
    cudaMemcpy(src1, h, n * 8, cudaMemcpyHostToDevice);
    cudaDeviceSynchronize();
    cudaSetDevice(0);
    cudaError_t e = cudaDeviceEnablePeerAccess(1, 0);
    printf("enable peer: %s\n", cudaGetErrorString(e));
    cudaMalloc(&out0, n * 8);
    cudaMemset(out0, 0, n * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8 + 64);
    k<<<1, 256, n * 8 + 64>>>(src1, out0, n);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    double *r = new double[n];
    cudaMemcpy(r, out0, n * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < n; ++i) bad += r[i] != h[i];
    printf("TMA bulk copy from peer memory: %d mismatches of %d (first values %g %g)\n", bad, n, r[0], r[1]);
    return 0;
}
