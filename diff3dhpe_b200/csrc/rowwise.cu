// Row-wise (memory-bound) kernels: one warp owns one 512-wide token row, lane l holds columns
// {128*i + 4*l .. +3 : i = 0..3} so every global access is a fully coalesced 128-bit transaction.
//
//   lift_ln            G-lift : fusion_layer (K=5) + Spatial_pos_embed + block-0 time vector + norm1   (MODEL:250,230-233,113-116,127)
//   postnorm_add_ln    G-ln   : Spatial_/Temporal_norm + Temporal_pos_embed + next block's time vector + next norm1
//   ln_split           G-ln   : norm2 -> split fp16 A operand of fc1                                  (MODEL:128)
//   head_ddim          G-head+ddim : Temporal_norm + head LayerNorm(1e-5) + Linear(512->3) + clamp + DDIM update
//                                                                                  (MODEL:245,255; DIFF:256,283-297)
//   tta_merge, mpjpe   G-tta  : RUN:583-588, LOSS:15-27
#include "kernels.cuh"
#include "operand.cuh"

namespace d3d {
namespace {

constexpr int kWarpsPerCta = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void load_row(const float* __restrict__ p, int lane, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + 128 * i + 4 * lane);
    v[4 * i + 0] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void load_row_ldg(const float* __restrict__ p, int lane, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + 128 * i + 4 * lane));
    v[4 * i + 0] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void store_row(float* __restrict__ p, int lane, const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(p + 128 * i + 4 * lane) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// GEMM A operand of the row (K = 512) in the handle's operand format (operand.cuh)
template <int FMT>
__device__ __forceinline__ void store_operand(__half* __restrict__ hi, __half* __restrict__ second, int lane,
                                              const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float x[4] = {v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]};
    op_store4<FMT>(hi, second, kC, 128 * i + 4 * lane, x);
  }
}

// y = (x - mean) * rstd * gamma + beta over 512 columns held by the warp (two-pass variance in registers).
__device__ __forceinline__ void layernorm_row(const float (&x)[16], const float* __restrict__ gamma,
                                              const float* __restrict__ beta, float eps, int lane, float (&y)[16]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  const float mean = warp_sum(s) * (1.0f / kC);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) { const float d = x[i] - mean; q = fmaf(d, d, q); }
  const float var = warp_sum(q) * (1.0f / kC);
  const float rstd = 1.0f / sqrtf(var + eps);
  float g[16], b[16];
  load_row_ldg(gamma, lane, g);
  load_row_ldg(beta, lane, b);
#pragma unroll
  for (int i = 0; i < 16; ++i) y[i] = fmaf((x[i] - mean) * rstd, g[i], b[i]);
}

template <int FMT>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
lift_ln_kernel(const float* __restrict__ x2d, const float* __restrict__ y3, const float* __restrict__ x5,
               const float* __restrict__ wf_t, const float* __restrict__ bf, const float* __restrict__ spos,
               const float* __restrict__ tvec, int64_t tvec_stride, LnParams ln1, float* __restrict__ X,
               __half* __restrict__ a_hi, __half* __restrict__ a_lo, int64_t T, int J, int tokens_per_clip) {
  const int lane = threadIdx.x & 31;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (t >= T) return;
  float in[5];
  if (x5) {
#pragma unroll
    for (int k = 0; k < 5; ++k) in[k] = __ldg(x5 + t * 5 + k);
  } else {
    in[0] = __ldg(x2d + t * 2); in[1] = __ldg(x2d + t * 2 + 1);
    in[2] = __ldg(y3 + t * 3); in[3] = __ldg(y3 + t * 3 + 1); in[4] = __ldg(y3 + t * 3 + 2);
  }
  float v[16], w[16];
  // F.linear: sum_k in_k * W[c][k] + b[c]
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    load_row_ldg(wf_t + k * kC, lane, w);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaf(in[k], w[i], v[i]);
  }
  load_row_ldg(bf, lane, w);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += w[i];
  load_row_ldg(spos + static_cast<int64_t>(t % J) * kC, lane, w);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += w[i];
  if (tvec) {
    load_row_ldg(tvec + (t / tokens_per_clip) * tvec_stride, lane, w);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += w[i];
  }
  store_row(X + t * kC, lane, v);
  float a[16];
  layernorm_row(v, ln1.gamma, ln1.beta, 1e-6f, lane, a);
  store_operand<FMT>(a_hi + t * kC, a_lo + t * kC, lane, a);
}

template <int FMT>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
postnorm_add_ln_kernel(float* __restrict__ X, LnParams post, const float* __restrict__ tpos,
                       const float* __restrict__ tvec, int64_t tvec_stride, LnParams ln1, __half* __restrict__ a_hi,
                       __half* __restrict__ a_lo, int64_t T, int J, int F) {
  const int lane = threadIdx.x & 31;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (t >= T) return;
  float x[16], z[16], w[16];
  load_row(X + t * kC, lane, x);
  layernorm_row(x, post.gamma, post.beta, 1e-6f, lane, z);
  if (tpos) {
    load_row_ldg(tpos + ((t / J) % F) * kC, lane, w);
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] += w[i];
  }
  if (tvec) {
    load_row_ldg(tvec + (t / (static_cast<int64_t>(J) * F)) * tvec_stride, lane, w);
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] += w[i];
  }
  store_row(X + t * kC, lane, z);
  layernorm_row(z, ln1.gamma, ln1.beta, 1e-6f, lane, x);
  store_operand<FMT>(a_hi + t * kC, a_lo + t * kC, lane, x);
}

template <int FMT>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
ln_split_kernel(const float* __restrict__ X, LnParams ln, float eps, __half* __restrict__ a_hi,
                __half* __restrict__ a_lo, int64_t T) {
  const int lane = threadIdx.x & 31;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (t >= T) return;
  float x[16], y[16];
  load_row(X + t * kC, lane, x);
  layernorm_row(x, ln.gamma, ln.beta, eps, lane, y);
  store_operand<FMT>(a_hi + t * kC, a_lo + t * kC, lane, y);
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
ln_f32_kernel(const float* __restrict__ X, LnParams ln, float eps, float* __restrict__ out, int64_t T) {
  const int lane = threadIdx.x & 31;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (t >= T) return;
  float x[16], y[16];
  load_row(X + t * kC, lane, x);
  layernorm_row(x, ln.gamma, ln.beta, eps, lane, y);
  store_row(out + t * kC, lane, y);
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
head_ddim_kernel(const float* __restrict__ X, LnParams post, LnParams head_ln, const float* __restrict__ wh,
                 const float* __restrict__ bh, DdimStep s, float* __restrict__ y, const float* __restrict__ noise,
                 float* __restrict__ out3, float* __restrict__ trace_y, float* __restrict__ trace_x0,
                 int trace_stride, int trace_idx, int64_t T) {
  const int lane = threadIdx.x & 31;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (t >= T) return;
  float x[16], z[16], h[16], w[16];
  load_row(X + t * kC, lane, x);
  layernorm_row(x, post.gamma, post.beta, 1e-6f, lane, z);
  layernorm_row(z, head_ln.gamma, head_ln.beta, 1e-5f, lane, h);
  float o[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    load_row_ldg(wh + k * kC, lane, w);
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) d = fmaf(h[i], w[i], d);
    o[k] = warp_sum(d) + __ldg(bh + k);
  }
  if (lane < 3) {
    float x0 = lane == 0 ? o[0] : (lane == 1 ? o[1] : o[2]);
    const int64_t e = t * 3 + lane;
    if (out3) {                     // forward_denoise: raw head output (MODEL:255-257)
      out3[e] = x0;
      return;
    }
    if (s.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);          // DIFF:252,256
    float yn;
    if (s.last) {
      yn = x0;                                               // DIFF:283-285
    } else {
      // literal DIFF:295-297, one rounding per torch op (no FMA contraction)
      const float t1 = __fmul_rn(x0, s.sqrt_alpha_next);
      const float t2 = __fdiv_rn(__fsub_rn(y[e], __fmul_rn(s.alpha, x0)), s.sqrt_one_minus);
      yn = __fadd_rn(t1, __fmul_rn(s.c, t2));
      if (noise) yn = __fadd_rn(yn, __fmul_rn(s.sigma, noise[e]));
    }
    y[e] = yn;
    if (trace_y) trace_y[e * trace_stride + trace_idx] = yn;
    if (trace_x0) trace_x0[e * trace_stride + trace_idx] = x0;
  }
}

__global__ void tta_merge_kernel(const float* __restrict__ y, const float* __restrict__ yf,
                                 const int32_t* __restrict__ perm, float scale, float* __restrict__ out,
                                 int64_t n_elems, int J) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_elems) return;
  const int k = static_cast<int>(e % 3);
  const int64_t fj = e / 3;
  const int j = static_cast<int>(fj % J);
  const int64_t f = fj / J;
  float v = yf[(f * J + perm[j]) * 3 + k];
  if (k == 0) v = -v;
  out[e] = __fmul_rn(__fdiv_rn(__fadd_rn(y[e], v), 2.0f), scale);     // RUN:587-588
}

__global__ void mpjpe_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                             const uint8_t* __restrict__ mask, int64_t n_joints_total, int J,
                             double* __restrict__ acc) {
  double s = 0.0, c = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_joints_total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (mask && !mask[i / J]) continue;
    const float dx = pred[i * 3] - gt[i * 3], dy = pred[i * 3 + 1] - gt[i * 3 + 1], dz = pred[i * 3 + 2] - gt[i * 3 + 2];
    s += static_cast<double>(sqrtf(dx * dx + dy * dy + dz * dz));
    c += 1.0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  __shared__ double ss[32], cs[32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { ss[w] = s; cs[w] = c; }
  __syncthreads();
  if (w == 0) {
    s = l < (blockDim.x >> 5) ? ss[l] : 0.0;
    c = l < (blockDim.x >> 5) ? cs[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (l == 0) { atomicAdd(acc, s); atomicAdd(acc + 1, c); }
  }
}

// fp32 [rows, K] -> operand arrays.  is_weight selects the weight-side e5m2 scales of FMT_F8C.
__global__ void split_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ second,
                             int64_t n, int K, int fmt, int is_weight, float* __restrict__ absmax) {
  float amax = 0.f;
  bool bad = false;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = in[i];
    if (isfinite(v)) amax = fmaxf(amax, fabsf(v)); else bad = true;
    const __half h = __float2half_rn(v);
    const float l = v - __half2float(h);
    hi[i] = h;
    if (!second) continue;
    if (fmt == FMT_SPLIT16) {
      second[i] = __float2half_rn(l);
    } else {
      const int64_t row = i / K;
      const int col = static_cast<int>(i - row * K);
      uint8_t* c8 = reinterpret_cast<uint8_t*>(second) + row * 2 * K;
      const float first = is_weight ? l * kWgtLoScale : v * kActHiScale;
      const float secnd = is_weight ? v * kWgtHiScale : l * kActLoScale;
      c8[col] = static_cast<uint8_t>(op_e5m2x2(first, 0.f) & 0xff);
      c8[K + col] = static_cast<uint8_t>(op_e5m2x2(secnd, 0.f) & 0xff);
    }
  }
  if (absmax) {       // non-negative floats order like their bit patterns
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(absmax), __float_as_uint(amax));
    if (bad) absmax[1] = 1.0f;
  }
}
// activation operand -> fp32 (hi + lo); for FMT_F8C the lo term comes back from its e5m2 image
__global__ void merge_kernel(const __half* __restrict__ hi, const __half* __restrict__ second, float* __restrict__ out,
                             int64_t n, int K, int fmt) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float lo;
    if (fmt == FMT_SPLIT16) {
      lo = __half2float(second[i]);
    } else {
      const int64_t row = i / K;
      const int col = static_cast<int>(i - row * K);
      lo = op_e5m2_to_float(reinterpret_cast<const uint8_t*>(second)[row * 2 * K + K + col]) * (1.0f / kActLoScale);
    }
    out[i] = __half2float(hi[i]) + lo;
  }
}

inline unsigned row_grid(int64_t T) { return static_cast<unsigned>((T + kWarpsPerCta - 1) / kWarpsPerCta); }
inline unsigned flat_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  return static_cast<unsigned>(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g));
}

}  // namespace

cudaError_t launch_split(const float* in, __half* hi, __half* second, int64_t rows, int K, int fmt, int is_weight,
                         cudaStream_t st, float* absmax) {
  const int64_t n = rows * K;
  if (n <= 0) return cudaSuccess;
  split_kernel<<<flat_grid(n), 256, 0, st>>>(in, hi, second, n, K, fmt, is_weight, absmax);
  return cudaGetLastError();
}
cudaError_t launch_merge(const __half* hi, const __half* second, float* out, int64_t rows, int K, int fmt,
                         cudaStream_t st) {
  const int64_t n = rows * K;
  if (n <= 0) return cudaSuccess;
  merge_kernel<<<flat_grid(n), 256, 0, st>>>(hi, second, out, n, K, fmt);
  return cudaGetLastError();
}
cudaError_t launch_lift_ln(const float* x2d, const float* y, const float* x5, const float* wf_t, const float* bf,
                           const float* spos, const float* tvec, int64_t tvec_stride, LnParams ln1, float* X,
                           __half* a_hi, __half* a_lo, int fmt, int64_t T, int J, int tokens_per_clip,
                           cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  auto kern = fmt == FMT_F8C ? lift_ln_kernel<FMT_F8C> : lift_ln_kernel<FMT_SPLIT16>;
  kern<<<row_grid(T), kWarpsPerCta * 32, 0, st>>>(x2d, y, x5, wf_t, bf, spos, tvec, tvec_stride, ln1, X, a_hi, a_lo, T,
                                                 J, tokens_per_clip);
  return cudaGetLastError();
}
cudaError_t launch_postnorm_add_ln(float* X, LnParams post, const float* tpos, const float* tvec,
                                   int64_t tvec_stride, LnParams ln1, __half* a_hi, __half* a_lo, int fmt, int64_t T,
                                   int J, int F, cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  auto kern = fmt == FMT_F8C ? postnorm_add_ln_kernel<FMT_F8C> : postnorm_add_ln_kernel<FMT_SPLIT16>;
  kern<<<row_grid(T), kWarpsPerCta * 32, 0, st>>>(X, post, tpos, tvec, tvec_stride, ln1, a_hi, a_lo, T, J, F);
  return cudaGetLastError();
}
cudaError_t launch_ln_split(const float* X, LnParams ln, float eps, __half* a_hi, __half* a_lo, int fmt, int64_t T,
                            cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  auto kern = fmt == FMT_F8C ? ln_split_kernel<FMT_F8C> : ln_split_kernel<FMT_SPLIT16>;
  kern<<<row_grid(T), kWarpsPerCta * 32, 0, st>>>(X, ln, eps, a_hi, a_lo, T);
  return cudaGetLastError();
}
cudaError_t launch_ln_f32(const float* x, LnParams ln, float eps, float* out, int64_t T, cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  ln_f32_kernel<<<row_grid(T), kWarpsPerCta * 32, 0, st>>>(x, ln, eps, out, T);
  return cudaGetLastError();
}
cudaError_t launch_head_ddim(const float* X, LnParams post, LnParams head_ln, const float* wh, const float* bh,
                             DdimStep s, float* y, const float* noise, float* out3, float* trace_y, float* trace_x0,
                             int trace_stride, int trace_idx, int64_t T, cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  head_ddim_kernel<<<row_grid(T), kWarpsPerCta * 32, 0, st>>>(X, post, head_ln, wh, bh, s, y, noise, out3, trace_y,
                                                             trace_x0, trace_stride, trace_idx, T);
  return cudaGetLastError();
}
cudaError_t launch_tta_merge(const float* y, const float* yf, const int32_t* perm, float scale, float* out,
                             int64_t n_frames, int J, cudaStream_t st) {
  const int64_t n = n_frames * J * 3;
  if (n <= 0) return cudaSuccess;
  tta_merge_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(y, yf, perm, scale, out, n, J);
  return cudaGetLastError();
}
cudaError_t launch_mpjpe(const float* pred, const float* gt, const uint8_t* mask, int64_t n_frames, int J,
                         double* acc, cudaStream_t st) {
  const int64_t n = n_frames * J;
  if (n <= 0) return cudaSuccess;
  mpjpe_kernel<<<flat_grid(n), 256, 0, st>>>(pred, gt, mask, n, J, acc);
  return cudaGetLastError();
}

}  // namespace d3d
