// G-tattn: temporal GRAND attention (MODEL:76-83 with the '(b p) f c' grouping of MODEL:121) on tensor cores.
//
// One CTA per (clip b, joint j, head): the F frames of that joint are gathered straight out of the packed
// qkv tensor with a stride of J tokens (no transpose copies).  K and V of the sequence are split into fp16
// hi/lo halves and parked in shared memory once; each warp then owns 16-query tiles and runs a flash-style
// single sweep over 64-key chunks with mma.sync.m16n8k16 (fp32 accumulate):
//     S = Q_hi K_hi^T + Q_hi K_lo^T + Q_lo K_hi^T          (3-pass split, ~fp32 accurate)
//     online softmax in fp32 (scale 0.125), P split into hi/lo
//     O += P_hi V_hi + P_hi V_lo + P_lo V_hi
// and finally  out = O / l - V[query]   (GRAND: (P - I) V == P V - V).
// The sequence is at most 256 keys, so the whole K/V slab is smem resident (4 x NK x 144 B).
#include "kernels.cuh"

namespace d3d {
namespace {

constexpr int KS = 72;            // smem row stride in halves (144 B: ldmatrix rows hit distinct banks)
constexpr float kScale = 0.125f;

__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
  return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half hx = __float2half_rn(x), hy = __float2half_rn(y);
  hi = pack2(hx, hy);
  lo = pack2(__float2half_rn(x - __half2float(hx)), __float2half_rn(y - __half2float(hy)));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256)
attn_temporal_mma_kernel(const float* __restrict__ qkv, __half* __restrict__ o_hi, __half* __restrict__ o_lo,
                         float* __restrict__ o_f32, int F, int J, int NK) {
  extern __shared__ __align__(16) __half smh[];
  __half* Khi = smh;
  __half* Klo = Khi + NK * KS;
  __half* Vhi = Klo + NK * KS;
  __half* Vlo = Vhi + NK * KS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int head = blockIdx.y;
  const int64_t seq = blockIdx.x;                                   // b * J + j
  const int64_t tok0 = (seq / J) * (static_cast<int64_t>(F) * J) + (seq % J);
  const float* base = qkv + head * kHd;

  // ---- stage K, V (fp32 -> fp16 hi/lo), zero rows beyond F
  for (int i = threadIdx.x; i < NK * 16; i += blockDim.x) {
    const int r = i >> 4, c4 = i & 15;
    float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
    if (r < F) {
      const size_t row = static_cast<size_t>(tok0 + static_cast<int64_t>(r) * J) * (3 * kC);
      kk = *reinterpret_cast<const float4*>(base + row + kC + 4 * c4);
      vv = *reinterpret_cast<const float4*>(base + row + 2 * kC + 4 * c4);
    }
    uint32_t h0, l0, h1, l1;
    split2(kk.x, kk.y, h0, l0); split2(kk.z, kk.w, h1, l1);
    *reinterpret_cast<uint2*>(Khi + r * KS + 4 * c4) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(Klo + r * KS + 4 * c4) = make_uint2(l0, l1);
    split2(vv.x, vv.y, h0, l0); split2(vv.z, vv.w, h1, l1);
    *reinterpret_cast<uint2*>(Vhi + r * KS + 4 * c4) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(Vlo + r * KS + 4 * c4) = make_uint2(l0, l1);
  }
  __syncthreads();

  const uint32_t sKhi = static_cast<uint32_t>(__cvta_generic_to_shared(Khi));
  const uint32_t sKlo = static_cast<uint32_t>(__cvta_generic_to_shared(Klo));
  const uint32_t sVhi = static_cast<uint32_t>(__cvta_generic_to_shared(Vhi));
  const uint32_t sVlo = static_cast<uint32_t>(__cvta_generic_to_shared(Vlo));
  const int g = lane >> 2, q4 = lane & 3;
  // ldmatrix lane -> row/col offsets
  const int k_key = (lane & 7) + ((lane >> 4) << 3);     // K (B operand, non-trans): key offset within 16
  const int k_d = ((lane >> 3) & 1) << 3;                //                            d offset (0 / 8)
  const int v_key = (lane & 7) + (((lane >> 3) & 1) << 3);   // V (trans): key offset within 16
  const int v_d = (lane >> 4) << 3;                          //            d offset (0 / 8)

  const int n_qt = (F + 15) >> 4;
  const int n_chunks = NK >> 6;
  for (int qt = warp; qt < n_qt; qt += nwarps) {
    const int q0 = qt << 4;
    const int r0 = q0 + g, r1 = r0 + 8;
    // ---- Q fragments (A operand), split hi/lo
    uint32_t qh[4][4], ql[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c = ks * 16 + 2 * q4;
      float2 v00 = make_float2(0.f, 0.f), v10 = v00, v01 = v00, v11 = v00;
      if (r0 < F) {
        const float* p = base + static_cast<size_t>(tok0 + static_cast<int64_t>(r0) * J) * (3 * kC);
        v00 = *reinterpret_cast<const float2*>(p + c);
        v01 = *reinterpret_cast<const float2*>(p + c + 8);
      }
      if (r1 < F) {
        const float* p = base + static_cast<size_t>(tok0 + static_cast<int64_t>(r1) * J) * (3 * kC);
        v10 = *reinterpret_cast<const float2*>(p + c);
        v11 = *reinterpret_cast<const float2*>(p + c + 8);
      }
      split2(v00.x, v00.y, qh[ks][0], ql[ks][0]);
      split2(v10.x, v10.y, qh[ks][1], ql[ks][1]);
      split2(v01.x, v01.y, qh[ks][2], ql[ks][2]);
      split2(v11.x, v11.y, qh[ks][3], ql[ks][3]);
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }

    for (int kc = 0; kc < n_chunks; ++kc) {
      const int key0 = kc << 6;
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          const uint32_t off = static_cast<uint32_t>(((key0 + np * 16 + k_key) * KS + ks * 16 + k_d) * 2);
          uint32_t bh[4], bl[4];
          ldsm_x4(sKhi + off, bh);
          ldsm_x4(sKlo + off, bl);
          mma16816(s[2 * np], qh[ks], bh[0], bh[1]);
          mma16816(s[2 * np], qh[ks], bl[0], bl[1]);
          mma16816(s[2 * np], ql[ks], bh[0], bh[1]);
          mma16816(s[2 * np + 1], qh[ks], bh[2], bh[3]);
          mma16816(s[2 * np + 1], qh[ks], bl[2], bl[3]);
          mma16816(s[2 * np + 1], ql[ks], bh[2], bh[3]);
        }
      }
      // ---- scale, mask, online softmax
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = key0 + nt * 8 + 2 * q4 + (e & 1);
          const float v = key < F ? s[nt][e] * kScale : -INFINITY;
          s[nt][e] = v;
          if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);      // finite: every chunk holds >= 1 valid key
      const float c0 = expf(m0 - mn0), c1 = expf(m1 - mn1);        // exp(-inf) = 0 on the first chunk
      m0 = mn0; m1 = mn1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = expf(s[nt][0] - mn0); s[nt][1] = expf(s[nt][1] - mn0);
        s[nt][2] = expf(s[nt][2] - mn1); s[nt][3] = expf(s[nt][3] - mn1);
        rs0 += s[nt][0] + s[nt][1];
        rs1 += s[nt][2] + s[nt][3];
      }
      l0 = l0 * c0 + rs0;
      l1 = l1 * c1 + rs1;
#pragma unroll
      for (int i = 0; i < 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
      // ---- O += P V
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t ph[4], pl[4];
        split2(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
        split2(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
        split2(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
        split2(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          const uint32_t off = static_cast<uint32_t>(((key0 + kk * 16 + v_key) * KS + dp * 16 + v_d) * 2);
          uint32_t vh[4], vl[4];
          ldsm_x4_t(sVhi + off, vh);
          ldsm_x4_t(sVlo + off, vl);
          mma16816(o[2 * dp], ph, vh[0], vh[1]);
          mma16816(o[2 * dp], ph, vl[0], vl[1]);
          mma16816(o[2 * dp], pl, vh[0], vh[1]);
          mma16816(o[2 * dp + 1], ph, vh[2], vh[3]);
          mma16816(o[2 * dp + 1], ph, vl[2], vl[3]);
          mma16816(o[2 * dp + 1], pl, vh[2], vh[3]);
        }
      }
    }
    // ---- finalize: out = O / l - V[query]
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = half ? r1 : r0;
      if (r >= F) continue;
      const float inv = half ? i1 : i0;
      const size_t off = static_cast<size_t>(tok0 + static_cast<int64_t>(r) * J) * kC + head * kHd;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int d = nt * 8 + 2 * q4;
        const __half2 vh = *reinterpret_cast<const __half2*>(Vhi + r * KS + d);
        const __half2 vl = *reinterpret_cast<const __half2*>(Vlo + r * KS + d);
        const float x0 = o[nt][2 * half] * inv - (__low2float(vh) + __low2float(vl));
        const float x1 = o[nt][2 * half + 1] * inv - (__high2float(vh) + __high2float(vl));
        if (o_f32) {
          *reinterpret_cast<float2*>(o_f32 + off + d) = make_float2(x0, x1);
        } else {
          uint32_t hh, ll;
          split2(x0, x1, hh, ll);
          *reinterpret_cast<uint32_t*>(o_hi + off + d) = hh;
          *reinterpret_cast<uint32_t*>(o_lo + off + d) = ll;
        }
      }
    }
  }
}

}  // namespace

cudaError_t configure_attention_mma() {
  return cudaFuncSetAttribute(attn_temporal_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              4 * 256 * KS * static_cast<int>(sizeof(__half)));
}

cudaError_t launch_attn_temporal_mma(const float* qkv, __half* o_hi, __half* o_lo, float* o_f32, int B, int F, int J,
                                     cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  if (F < 1 || F > 256) return cudaErrorInvalidValue;
  const int NK = (F + 63) / 64 * 64;
  const int smem = 4 * NK * KS * static_cast<int>(sizeof(__half));
  const int n_qt = (F + 15) / 16;
  const int warps = n_qt < 8 ? n_qt : 8;
  dim3 grid(static_cast<unsigned>(B) * J, kHeads);
  attn_temporal_mma_kernel<<<grid, warps * 32, smem, st>>>(qkv, o_hi, o_lo, o_f32, F, J, NK);
  return cudaGetLastError();
}

}  // namespace d3d
