// Load-time folding of a LayerNorm into the Linear that consumes it (deferred norm2, kernels.cuh EPI_GELU_DLN):
//
//   Linear(LN(x)) = rstd_r (x . W'^T - mean_r s) + c,   W'[n,k] = W[n,k] gamma[k],  s_n = sum_k W'[n,k],
//                                                       c_n = b_n + sum_k W[n,k] beta[k]          (MODEL:128, 51)
//
// One block per output row n; the sums are accumulated in fp64 (they are subtracted from the accumulator of a 512-long
// dot product at run time, so they must not add an error of their own).
#include "kernels.cuh"

namespace d3d {

namespace {

__global__ void __launch_bounds__(128) fold_ln_linear_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ bias,
                                                             float* __restrict__ w_out, float* __restrict__ colsum,
                                                             float* __restrict__ cbias, int K) {
  const int n = blockIdx.x;
  double s = 0.0, c = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float wv = w[static_cast<size_t>(n) * K + k];
    const float wp = wv * gamma[k];                 // the fp32 value the operand split sees
    w_out[static_cast<size_t>(n) * K + k] = wp;
    s += static_cast<double>(wp);
    c += static_cast<double>(wv) * static_cast<double>(beta[k]);
  }
  __shared__ double sh[2][4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    colsum[n] = static_cast<float>(sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3]);
    cbias[n] = static_cast<float>(static_cast<double>(bias[n]) + sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3]);
  }
}

}  // namespace

cudaError_t launch_fold_ln_linear(const float* w, const float* gamma, const float* beta, const float* bias, float* w_out,
                                  float* colsum, float* cbias, int N, int K, cudaStream_t st) {
  fold_ln_linear_kernel<<<N, 128, 0, st>>>(w, gamma, beta, bias, w_out, colsum, cbias, K);
  return cudaGetLastError();
}

}  // namespace d3d
