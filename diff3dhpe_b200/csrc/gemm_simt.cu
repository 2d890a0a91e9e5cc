// CUDA-core fp32 validation GEMM.  Consumes exactly the operands of the tcgen05 kernel (fp16 hi/lo halves,
// K-major) and produces exactly its outputs, so the tensor-core path can be checked element by element on
// the device.  Not a performance path: 64x64 tiles, 4x4 outputs per thread.
#include "kernels.cuh"

namespace d3d {
namespace {

constexpr int TM = 64, TN = 64, TK = 32;

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

template <int EPI>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const __half* __restrict__ a_hi, const __half* __restrict__ a_lo, const __half* __restrict__ b_hi,
                 const __half* __restrict__ b_lo, const GemmParams p) {
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += TK) {
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      const int r = i / TK, c = i % TK;
      const int gm = m0 + r;
      float v = 0.f;
      if (gm < p.M) {
        const size_t o = static_cast<size_t>(gm) * p.K + k0 + c;
        v = __half2float(a_hi[o]) + (a_lo ? __half2float(a_lo[o]) : 0.f);
      }
      As[c][r] = v;
      const size_t ob = static_cast<size_t>(n0 + r) * p.K + k0 + c;
      Bs[c][r] = __half2float(b_hi[ob]) + (b_lo ? __half2float(b_lo[ob]) : 0.f);
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      const size_t o = static_cast<size_t>(gm) * p.N + gn;
      float v = acc[i][j] + p.bias[gn];
      if (EPI == EPI_F32) {
        if (p.residual) v += p.residual[o];
        p.out_f32[o] = v;
      } else if (EPI == EPI_GELU_SPLIT) {
        v = gelu_erf(v);
        const __half h = __float2half_rn(v);
        p.out_hi[o] = h;
        p.out_lo[o] = __float2half_rn(v - __half2float(h));
      } else {   // EPI_QKV16
        const size_t oq = static_cast<size_t>(gm) * kQkvRow + gn;
        const __half h = __float2half_rn(v);
        p.out_qkv[oq] = h;
        if (gn >= 2 * kC) p.out_qkv[oq + kC] = __float2half_rn(v - __half2float(h));
      }
    }
  }
}

}  // namespace

cudaError_t launch_gemm_simt(const __half* a_hi, const __half* a_lo, const __half* b_hi, const __half* b_lo,
                             const GemmParams& p, int epi, cudaStream_t st) {
  if (p.M <= 0) return cudaSuccess;
  if (p.N % TN != 0 || p.K % TK != 0) return cudaErrorInvalidValue;
  dim3 grid((p.M + TM - 1) / TM, p.N / TN);
  if (epi == EPI_F32)
    gemm_simt_kernel<EPI_F32><<<grid, 256, 0, st>>>(a_hi, a_lo, b_hi, b_lo, p);
  else if (epi == EPI_GELU_SPLIT)
    gemm_simt_kernel<EPI_GELU_SPLIT><<<grid, 256, 0, st>>>(a_hi, a_lo, b_hi, b_lo, p);
  else
    gemm_simt_kernel<EPI_QKV16><<<grid, 256, 0, st>>>(a_hi, a_lo, b_hi, b_lo, p);
  return cudaGetLastError();
}

}  // namespace d3d
