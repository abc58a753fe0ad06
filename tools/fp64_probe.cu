// fp64_probe.cu — dependent-issue latency and single-warp throughput of DFMA on B200 (one warp per SM sub-partition)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b)
{
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) v[i] = fma(v[i], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
}
template <int ILP> void run(double* out, long long* cyc, int warps)
{
    long long h[148 * 32];
    const int iters = 4000;
    k<ILP><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(long long) * 148 * warps, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < 148 * warps; i++) mx = h[i] > mx ? h[i] : mx;
    printf("warps/SM %2d ILP %d: %.2f cycles per DFMA per warp\n", warps, ILP, mx / (iters * 8.0 * ILP));
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 32 * 8);
    for (int w : {1, 4, 8, 16}) { run<1>(out, cyc, w); run<2>(out, cyc, w); run<4>(out, cyc, w); run<8>(out, cyc, w); }
    return 0;
}
