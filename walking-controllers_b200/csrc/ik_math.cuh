// ik_math.cuh — small 3x3 / rigid-transform / warp helpers shared by the IK kernels (ik_kernels.cu, ik_tpi.cu).
#pragma once

#include <cuda_runtime.h>

namespace wcb {
namespace {

// ---- small 3x3 / rigid transform helpers (row-major R[9], p[3]) ---------------------------------
__device__ __forceinline__ void mat3_mul(const double* a, const double* b, double* o)
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) o[3 * r + c] = a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c] + a[3 * r + 2] * b[6 + c];
}
__device__ __forceinline__ void mat3_mul_bt(const double* a, const double* b, double* o)
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) o[3 * r + c] = a[3 * r] * b[3 * c] + a[3 * r + 1] * b[3 * c + 1] + a[3 * r + 2] * b[3 * c + 2];
}
__device__ __forceinline__ void xf_mul(const double* Ra, const double* pa, const double* Rb, const double* pb, double* Ro, double* po)
{
    double R[9], p[3];
    mat3_mul(Ra, Rb, R);
#pragma unroll
    for (int r = 0; r < 3; r++) p[r] = Ra[3 * r] * pb[0] + Ra[3 * r + 1] * pb[1] + Ra[3 * r + 2] * pb[2] + pa[r];
#pragma unroll
    for (int k = 0; k < 9; k++) Ro[k] = R[k];
#pragma unroll
    for (int k = 0; k < 3; k++) po[k] = p[k];
}
__device__ __forceinline__ void cross3(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
// unskew(1/2 (M - M')), M = R Rdes'   (Utils.cpp:22-27 + iDynTree::unskew)
__device__ __forceinline__ void rot_error(const double* R, const double* Rdes, double* e)
{
    double M[9];
    mat3_mul_bt(R, Rdes, M);
    e[0] = 0.5 * (M[7] - M[5]);
    e[1] = 0.5 * (M[2] - M[6]);
    e[2] = 0.5 * (M[3] - M[1]);
}
// 8-byte asynchronous copy HBM -> shared (LDGSTS): the per-instance inputs are fetched while the kinematics runs
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

} // namespace
} // namespace wcb
