// tmem_probe.cu — microbenchmark: can tensor memory serve as per-lane read-only scratch for fp64 data?
// Measures tcgen05.ld.32x32b.x2 (one double per lane) latency and throughput, 1..4 warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int BATCH>
__global__ void probe(double* out, long long* cyc, int iters, int mode)
{
    __shared__ uint32_t tbase_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&tbase_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tbase_s + ((uint32_t)(32 * (warp & 3)) << 16);
    // fill: column pair c holds double (lane + c)
    for (int c = 0; c < 256; c++) {
        const double v = (double)(lane + c);
        tmem_st2(tb + 2 * c, (uint32_t)__double_as_longlong(v), (uint32_t)(__double_as_longlong(v) >> 32));
    }
    tmem_wait_st();
    __syncwarp();
    double acc = 0.0;
    long long t0 = clock64();
    if (mode == 0) { // throughput: BATCH independent loads, one wait
        for (int it = 0; it < iters; it++) {
            uint32_t lo[BATCH], hi[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; b++) tmem_ld2(tb + 2 * ((it * BATCH + b) & 255), lo[b], hi[b]);
            tmem_wait_ld();
#pragma unroll
            for (int b = 0; b < BATCH; b++) acc += __longlong_as_double(((long long)hi[b] << 32) | lo[b]);
        }
    } else { // latency: each address depends on the previous value
        uint32_t col = 0;
        for (int it = 0; it < iters * BATCH; it++) {
            uint32_t lo, hi;
            tmem_ld2(tb + 2 * (col & 255), lo, hi);
            tmem_wait_ld();
            const double v = __longlong_as_double(((long long)hi << 32) | lo);
            acc += v;
            col = (uint32_t)(v - lane) + 1; // = previous column + 1, but data dependent
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase_s), "r"(512));
}

// wide loads: one instruction fetches W consecutive 32-bit columns per lane (W/2 doubles)
template <int W> __device__ __forceinline__ void tmem_ldw(uint32_t taddr, uint32_t* r);
template <> __device__ __forceinline__ void tmem_ldw<4>(uint32_t t, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(t));
}
template <> __device__ __forceinline__ void tmem_ldw<8>(uint32_t t, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t));
}
template <> __device__ __forceinline__ void tmem_ldw<16>(uint32_t t, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(t));
}

template <int W, int BATCH>
__global__ void probe_wide(double* out, long long* cyc, int iters)
{
    __shared__ uint32_t tbase_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&tbase_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tbase_s + ((uint32_t)(32 * (warp & 3)) << 16);
    for (int c = 0; c < 256; c++) {
        const double v = (double)(lane + c);
        tmem_st2(tb + 2 * c, (uint32_t)__double_as_longlong(v), (uint32_t)(__double_as_longlong(v) >> 32));
    }
    tmem_wait_st();
    __syncwarp();
    double accs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        uint32_t r[BATCH][W];
#pragma unroll
        for (int b = 0; b < BATCH; b++) tmem_ldw<W>(tb + ((it * BATCH + b) * W) % 512, r[b]);
        tmem_wait_ld();
#pragma unroll
        for (int b = 0; b < BATCH; b++)
#pragma unroll
            for (int k = 0; k < W / 2; k++) accs[(b * (W / 2) + k) & 7] += __longlong_as_double(((long long)r[b][2 * k + 1] << 32) | r[b][2 * k]);
    }
    long long t1 = clock64();
    double acc = 0.0;
    for (int k = 0; k < 8; k++) acc += accs[k];
    if (lane == 0) cyc[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase_s), "r"(512));
}

template <int W, int BATCH>
void run_wide(double* out, long long* cyc, long long* h, int iters)
{
    for (int warps = 1; warps <= 4; warps *= 4) {
        probe_wide<W, BATCH><<<148, warps * 32>>>(out, cyc, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
        cudaMemcpy(h, cyc, sizeof(long long) * 148 * warps, cudaMemcpyDeviceToHost);
        double mx = 0; for (int k = 0; k < 148 * warps; k++) mx = h[k] > mx ? h[k] : mx;
        printf("tmem  warps/SM %d x%d batch %d: %.2f cycles per instruction, %.2f cycles per double per lane\n", warps, W, BATCH,
               mx / (iters * (double)BATCH), mx / (iters * (double)BATCH * (W / 2)));
    }
}

// same access pattern from shared memory for comparison
template <int BATCH>
__global__ void probe_smem(double* out, long long* cyc, int iters)
{
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* base = sm + warp * 256 * 32 + lane;
    for (int c = 0; c < 256; c++) base[c * 32] = (double)(lane + c);
    __syncwarp();
    double acc = 0.0, accs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        double v[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; b++) v[b] = *(volatile double*)&base[((it * BATCH + b) & 255) * 32];
#pragma unroll
        for (int b = 0; b < BATCH; b++) accs[b & 7] += v[b];
    }
    long long t1 = clock64();
    for (int k = 0; k < 8; k++) acc += accs[k];
    if (lane == 0) cyc[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 128 * sizeof(double)); cudaMalloc(&cyc, 148 * 4 * sizeof(long long));
    long long h[148 * 4];
    const int iters = 2000;
    run_wide<4, 4>(out, cyc, h, iters); run_wide<8, 4>(out, cyc, h, iters); run_wide<16, 2>(out, cyc, h, iters);
    run_wide<8, 1>(out, cyc, h, iters);
    for (int warps = 1; warps <= 4; warps *= 2) {
        for (int mode = 0; mode < 2; mode++) {
            probe<16><<<148, warps * 32>>>(out, cyc, iters, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, cyc, sizeof(long long) * 148 * warps, cudaMemcpyDeviceToHost);
            double mx = 0; for (int k = 0; k < 148 * warps; k++) mx = h[k] > mx ? h[k] : mx;
            printf("tmem  warps/SM %d %s: %.2f cycles per ld.32x32b.x2 (max over warps)\n", warps, mode ? "dependent " : "batch of 16", mx / (iters * 16.0));
        }
        cudaFuncSetAttribute(probe_smem<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 256 * 32 * 8);
        if (warps == 4) continue;
        probe_smem<16><<<148, warps * 32, warps * 256 * 32 * 8>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(long long) * 148 * warps, cudaMemcpyDeviceToHost);
        double mx = 0; for (int k = 0; k < 148 * warps; k++) mx = h[k] > mx ? h[k] : mx;
        printf("smem  warps/SM %d batch of 16: %.2f cycles per LDS.64\n", warps, mx / (iters * 16.0));
    }
    return 0;
}
