// GRAND attention cores (MODEL:76-83):  O = (softmax(Q K^T * hd^-0.5) - I) V  on the packed qkv tensor
// [T, 3*512] written by the qkv GEMM (channel = which*512 + head*64 + d; token = (b*F + f)*J + j).
//
//   attn_spatial_kernel   G-sattn: the 17 joints of one frame.  One warp per (frame, head); K and V live in
//                         shared memory, each of the first 17 lanes owns one query row in registers, so the
//                         17x17 softmax needs no shuffles and P never leaves registers.
//   attn_generic_kernel   CUDA-core validation kernel for any sequence (temporal: the F frames of one joint,
//                         addressed with stride J tokens -- no transposes, MODEL:119-121,130-133 eliminated).
//   (the tensor-core temporal kernel lives in attention_mma.cu)
//
// Output is written token-major [T, 512] (heads merged, MODEL:83) either as the split-fp16 A operand of the
// proj GEMM or as fp32 (op-level tests).
#include "kernels.cuh"

namespace d3d {
namespace {

constexpr float kScale = 0.125f;   // head_dim ** -0.5, MODEL:65

__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
  return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}

// ------------------------------------------------------------------------------------------ spatial, J = 17
constexpr int SJ = 17;

__global__ void __launch_bounds__(256)
attn_spatial_kernel(const float* __restrict__ qkv, __half* __restrict__ o_hi, __half* __restrict__ o_lo,
                    float* __restrict__ o_f32, int64_t n_groups) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = warp;                                    // 8 warps = 8 heads of one frame
  const int64_t g = blockIdx.x;
  if (g >= n_groups) return;
  float* Ks = sm + warp * (2 * SJ * kHd);
  float* Vs = Ks + SJ * kHd;
  const float* base = qkv + g * SJ * (3 * kC) + head * kHd;

  // K, V rows -> smem (17 rows x 16 float4 each)
  for (int i = lane; i < SJ * 16; i += 32) {
    const int r = i >> 4, c4 = i & 15;
    const float4 kk = *reinterpret_cast<const float4*>(base + static_cast<size_t>(r) * (3 * kC) + kC + 4 * c4);
    const float4 vv = *reinterpret_cast<const float4*>(base + static_cast<size_t>(r) * (3 * kC) + 2 * kC + 4 * c4);
    *reinterpret_cast<float4*>(Ks + r * kHd + 4 * c4) = kk;
    *reinterpret_cast<float4*>(Vs + r * kHd + 4 * c4) = vv;
  }
  // own query row -> registers
  const int qi = lane < SJ ? lane : SJ - 1;                 // idle lanes shadow row 16 (results discarded)
  float q[kHd];
#pragma unroll
  for (int c4 = 0; c4 < 16; ++c4) {
    const float4 t = *reinterpret_cast<const float4*>(base + static_cast<size_t>(qi) * (3 * kC) + 4 * c4);
    q[4 * c4] = t.x; q[4 * c4 + 1] = t.y; q[4 * c4 + 2] = t.z; q[4 * c4 + 3] = t.w;
  }
  __syncwarp();

  float s[SJ];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < SJ; ++j) {
    float a = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
      const float4 kk = *reinterpret_cast<const float4*>(Ks + j * kHd + 4 * c4);
      a = fmaf(q[4 * c4], kk.x, a); a = fmaf(q[4 * c4 + 1], kk.y, a);
      a = fmaf(q[4 * c4 + 2], kk.z, a); a = fmaf(q[4 * c4 + 3], kk.w, a);
    }
    s[j] = a * kScale;
    mx = fmaxf(mx, s[j]);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < SJ; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int j = 0; j < SJ; ++j) s[j] = s[j] * inv - (j == qi ? 1.0f : 0.0f);      // P - I  (MODEL:82-83)

  float o[kHd];
#pragma unroll
  for (int d = 0; d < kHd; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < SJ; ++j) {
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
      const float4 vv = *reinterpret_cast<const float4*>(Vs + j * kHd + 4 * c4);
      o[4 * c4] = fmaf(s[j], vv.x, o[4 * c4]); o[4 * c4 + 1] = fmaf(s[j], vv.y, o[4 * c4 + 1]);
      o[4 * c4 + 2] = fmaf(s[j], vv.z, o[4 * c4 + 2]); o[4 * c4 + 3] = fmaf(s[j], vv.w, o[4 * c4 + 3]);
    }
  }
  if (lane < SJ) {
    const size_t off = static_cast<size_t>(g * SJ + lane) * kC + head * kHd;
    if (o_f32) {
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4)
        *reinterpret_cast<float4*>(o_f32 + off + 4 * c4) = make_float4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
    } else {
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v0 = o[8 * c8 + 2 * e], v1 = o[8 * c8 + 2 * e + 1];
          const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
          hw[e] = pack2(h0, h1);
          lw[e] = pack2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
        }
        *reinterpret_cast<uint4*>(o_hi + off + 8 * c8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(o_lo + off + 8 * c8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ generic (validation)
// One CTA per (sequence, head).  Token of position n in sequence s:  (s / inner) * outer + (s % inner) + n * tok_stride
//   spatial : inner = 1, outer = J, tok_stride = 1, N = J      (s = b*F + f)
//   temporal: inner = J, outer = F*J, tok_stride = J, N = F    (s = b*J + j)
constexpr int KPAD = kHd + 1;

__global__ void __launch_bounds__(256)
attn_generic_kernel(const float* __restrict__ qkv, __half* __restrict__ o_hi, __half* __restrict__ o_lo,
                    float* __restrict__ o_f32, int N, int64_t outer, int inner, int64_t tok_stride) {
  extern __shared__ float sm[];
  float* Ks = sm;                       // [N][65]
  float* Vs = Ks + ((N * KPAD + 3) & ~3);   // [N][64], 16-byte aligned
  float* Ps = Vs + N * kHd;             // [8][N]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int64_t s = blockIdx.x;
  const int64_t tok0 = (s / inner) * outer + (s % inner);
  const float* base = qkv + head * kHd;

  for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {
    const int r = i >> 4, c4 = i & 15;
    const size_t row = static_cast<size_t>(tok0 + r * tok_stride) * (3 * kC);
    const float4 kk = *reinterpret_cast<const float4*>(base + row + kC + 4 * c4);
    const float4 vv = *reinterpret_cast<const float4*>(base + row + 2 * kC + 4 * c4);
    float* kd = Ks + r * KPAD + 4 * c4;
    kd[0] = kk.x; kd[1] = kk.y; kd[2] = kk.z; kd[3] = kk.w;
    *reinterpret_cast<float4*>(Vs + r * kHd + 4 * c4) = vv;
  }
  __syncthreads();

  float* P = Ps + warp * N;
  for (int i = warp; i < N; i += 8) {
    float q[kHd];
    const float* qrow = base + static_cast<size_t>(tok0 + i * tok_stride) * (3 * kC);
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(qrow + 4 * c4));
      q[4 * c4] = t.x; q[4 * c4 + 1] = t.y; q[4 * c4 + 2] = t.z; q[4 * c4 + 3] = t.w;
    }
    float sc[8];
    float mx = -INFINITY;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int kk = lane + 32 * it;
      float a = -INFINITY;
      if (kk < N) {
        a = 0.f;
        const float* kr = Ks + kk * KPAD;
#pragma unroll
        for (int d = 0; d < kHd; ++d) a = fmaf(q[d], kr[d], a);
        a *= kScale;
      }
      sc[it] = a;
      mx = fmaxf(mx, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int kk = lane + 32 * it;
      sc[it] = kk < N ? expf(sc[it] - mx) : 0.f;
      sum += sc[it];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int kk = lane + 32 * it;
      if (kk < N) P[kk] = sc[it] * inv - (kk == i ? 1.0f : 0.0f);
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int kk = 0; kk < N; ++kk) {
      const float p = P[kk];
      const float2 vv = *reinterpret_cast<const float2*>(Vs + kk * kHd + 2 * lane);
      o0 = fmaf(p, vv.x, o0);
      o1 = fmaf(p, vv.y, o1);
    }
    __syncwarp();
    const size_t off = static_cast<size_t>(tok0 + i * tok_stride) * kC + head * kHd + 2 * lane;
    if (o_f32) {
      *reinterpret_cast<float2*>(o_f32 + off) = make_float2(o0, o1);
    } else {
      const __half h0 = __float2half_rn(o0), h1 = __float2half_rn(o1);
      *reinterpret_cast<uint32_t*>(o_hi + off) = pack2(h0, h1);
      *reinterpret_cast<uint32_t*>(o_lo + off) =
          pack2(__float2half_rn(o0 - __half2float(h0)), __float2half_rn(o1 - __half2float(h1)));
    }
  }
}

constexpr int kSpatialSmem = 8 * 2 * SJ * kHd * sizeof(float);   // 69632 B

}  // namespace

cudaError_t configure_attention() {
  cudaError_t e = cudaFuncSetAttribute(attn_spatial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpatialSmem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attn_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (((256 * KPAD + 3) & ~3) + 256 * kHd + 8 * 256) * static_cast<int>(sizeof(float)));
}

cudaError_t launch_attn_spatial(const float* qkv, __half* o_hi, __half* o_lo, float* o_f32, int64_t n_groups,
                                int J, cudaStream_t st) {
  if (n_groups <= 0) return cudaSuccess;
  if (J != SJ) return cudaErrorInvalidValue;
  const int smem = kSpatialSmem;
  attn_spatial_kernel<<<static_cast<unsigned>(n_groups), 256, smem, st>>>(qkv, o_hi, o_lo, o_f32, n_groups);
  return cudaGetLastError();
}

cudaError_t launch_attn_generic_simt(const float* qkv, __half* o_hi, __half* o_lo, float* o_f32, int n_seq, int N,
                                     int64_t outer, int inner, int64_t tok_stride, cudaStream_t st) {
  if (n_seq <= 0) return cudaSuccess;
  if (N < 1 || N > 256) return cudaErrorInvalidValue;
  const int smem = (((N * KPAD + 3) & ~3) + N * kHd + 8 * N) * static_cast<int>(sizeof(float));
  dim3 grid(static_cast<unsigned>(n_seq), kHeads);
  attn_generic_kernel<<<grid, 256, smem, st>>>(qkv, o_hi, o_lo, o_f32, N, outer, inner, tok_stride);
  return cudaGetLastError();
}

cudaError_t launch_attn_temporal_simt(const float* qkv, __half* o_hi, __half* o_lo, float* o_f32, int B, int F, int J,
                                      cudaStream_t st) {
  return launch_attn_generic_simt(qkv, o_hi, o_lo, o_f32, B * J, F, static_cast<int64_t>(F) * J, J, J, st);
}

}  // namespace d3d
