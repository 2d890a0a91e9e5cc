// G-gemm: the qkv / proj / fc1 / fc2 linears of MixSTE (MODEL:75, 84, 51-54) as a persistent,
// warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[M,N] = epilogue( A[M,K] . W[N,K]^T + bias[N] )
//
//   * operands are fp16 "split halves": x = hi + lo with hi = fp16(x), lo = fp16(x - hi).  The default
//     3-pass mode issues   D += A_hi.B_hi ; D += A_hi.B_lo ; D += A_lo.B_hi   per k-step, accumulating in
//     fp32 in TMEM, which restores ~22 mantissa bits (single-pass bf16/fp16 misses the
//     parity bar).  The 1-pass mode issues only A_hi.B_hi.
//   * TMA (cp.async.bulk.tensor, SWIZZLE_128B) stages {128 x 64} fp16 boxes; a ring of mbarrier-guarded
//     stages feeds one MMA-issuing thread (tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16).
//   * accumulators are double-buffered in TMEM (2 x BN fp32 columns) so the epilogue of tile i overlaps the
//     mainloop of tile i+1.  Eight epilogue warps read TMEM with tcgen05.ld (32 lanes x 32 columns) and apply
//     bias (+residual) -> fp32, or bias + exact-erf GELU -> split fp16 (the A operand of fc2).
//
// Warp roles (384 threads): 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = idle, 4..11 = epilogue.
#include "kernels.cuh"
#include "ptx.cuh"

namespace d3d {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                       // 64 fp16 = 128 B = one swizzle row
constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kTileBytesA = BM * BK * 2;     // 16 KiB
constexpr int kSmemBudget = 200 * 1024;

template <int BN, int PASSES>
struct Cfg {
  static constexpr int kTileBytesB = BN * BK * 2;
  static constexpr int kStageBytes = (PASSES == 3 ? 2 : 1) * (kTileBytesA + kTileBytesB);
  static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (kSmemBudget / kStageBytes);
  static constexpr int kTmemCols = 2 * BN;   // 256 or 512 (power of two)
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(kStages >= 2, "need at least a double buffer");
};

struct Barriers {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

template <int BN, int PASSES, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
               const GemmParams p) {
  using C = Cfg<BN, PASSES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + C::kStages * C::kStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_n = p.N / BN;
  const int n_tiles_m = (p.M + BM - 1) / BM;
  const int n_tiles = n_tiles_m * n_tiles_n;
  const int n_kb = p.K / BK;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tensormap(&tm_a_hi);
    ptx::prefetch_tensormap(&tm_b_hi);
    if (PASSES == 3) {
      ptx::prefetch_tensormap(&tm_a_lo);
      ptx::prefetch_tensormap(&tm_b_lo);
    }
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < C::kStages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&bars->tmem_full[a], 1);
      ptx::mbar_init(&bars->tmem_empty[a], kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<C::kTmemCols>(&bars->tmem_base);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles_n) * BM;
        const int n0 = (tile % n_tiles_n) * BN;
        for (int kb = 0; kb < n_kb; ++kb) {
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* s = smem + stage * C::kStageBytes;
          ptx::mbar_arrive_expect_tx(&bars->full[stage], C::kStageBytes);
          ptx::tma_load_2d(s, &tm_a_hi, &bars->full[stage], kb * BK, m0);
          s += kTileBytesA;
          if (PASSES == 3) {
            ptx::tma_load_2d(s, &tm_a_lo, &bars->full[stage], kb * BK, m0);
            s += kTileBytesA;
          }
#pragma unroll
          for (int h = 0; h < BN / 128; ++h)
            ptx::tma_load_2d(s + h * (128 * BK * 2), &tm_b_hi, &bars->full[stage], kb * BK, n0 + h * 128);
          s += C::kTileBytesB;
          if (PASSES == 3) {
#pragma unroll
            for (int h = 0; h < BN / 128; ++h)
              ptx::tma_load_2d(s + h * (128 * BK * 2), &tm_b_lo, &bars->full[stage], kb * BK, n0 + h * 128);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(BM, BN, 0 /*fp16*/);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < n_kb; ++kb) {
          ptx::mbar_wait(&bars->full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t s = ptx::smem_u32(smem + stage * C::kStageBytes);
          const uint32_t a_hi = s;
          const uint32_t a_lo = s + kTileBytesA;
          const uint32_t b_hi = s + (PASSES == 3 ? 2 : 1) * kTileBytesA;
          const uint32_t b_lo = b_hi + C::kTileBytesB;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da_hi = ptx::make_desc_k_sw128(a_hi + k * 32);
            const uint64_t db_hi = ptx::make_desc_k_sw128(b_hi + k * 32);
            ptx::mma_f16_ss(d_tmem, da_hi, db_hi, idesc, (kb | k) != 0 ? 1u : 0u);
            if (PASSES == 3) {
              const uint64_t da_lo = ptx::make_desc_k_sw128(a_lo + k * 32);
              const uint64_t db_lo = ptx::make_desc_k_sw128(b_lo + k * 32);
              ptx::mma_f16_ss(d_tmem, da_hi, db_lo, idesc, 1u);
              ptx::mma_f16_ss(d_tmem, da_lo, db_hi, idesc, 1u);
            }
          }
          ptx::mma_commit(&bars->empty[stage]);      // frees the smem stage once these MMAs have read it
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&bars->tmem_full[acc]);      // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (8 warps)
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;       // which half of the BN columns
    constexpr int kChunks = BN / 64;        // 32-column chunks per warp
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles_n) * BM;
      const int n0 = (tile % n_tiles_n) * BN;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      ptx::mbar_wait(&bars->tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ++ci) {
        const int col0 = (half * kChunks + ci) * 32;
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + acc * BN + col0 + (static_cast<uint32_t>(q * 32) << 16), r);
        ptx::tmem_ld_wait();
        const int gcol = n0 + col0;
        const float4* bias4 = reinterpret_cast<const float4*>(p.bias + gcol);
        if (EPI == EPI_F32) {
          float* orow = p.out_f32 + static_cast<size_t>(row) * p.N + gcol;
          const float* rrow = p.residual ? p.residual + static_cast<size_t>(row) * p.N + gcol : nullptr;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 b = __ldg(bias4 + v);
            float4 o;
            o.x = __uint_as_float(r[4 * v + 0]) + b.x;
            o.y = __uint_as_float(r[4 * v + 1]) + b.y;
            o.z = __uint_as_float(r[4 * v + 2]) + b.z;
            o.w = __uint_as_float(r[4 * v + 3]) + b.w;
            if (row_ok) {
              if (rrow) {
                const float4 rr = *reinterpret_cast<const float4*>(rrow + 4 * v);
                o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
              }
              *reinterpret_cast<float4*>(orow + 4 * v) = o;
            }
          }
        } else {
          __half* hrow = p.out_hi + static_cast<size_t>(row) * p.N + gcol;
          __half* lrow = p.out_lo + static_cast<size_t>(row) * p.N + gcol;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 b = __ldg(bias4 + 2 * v + (e >> 1));
              const float b0 = (e & 1) ? b.z : b.x;
              const float b1 = (e & 1) ? b.w : b.y;
              const float g0 = gelu_erf(__uint_as_float(r[8 * v + 2 * e + 0]) + b0);
              const float g1 = gelu_erf(__uint_as_float(r[8 * v + 2 * e + 1]) + b1);
              const __half h0 = __float2half_rn(g0), h1 = __float2half_rn(g1);
              const __half l0 = __float2half_rn(g0 - __half2float(h0));
              const __half l1 = __float2half_rn(g1 - __half2float(h1));
              hw[e] = static_cast<uint32_t>(__half_as_ushort(h0)) | (static_cast<uint32_t>(__half_as_ushort(h1)) << 16);
              lw[e] = static_cast<uint32_t>(__half_as_ushort(l0)) | (static_cast<uint32_t>(__half_as_ushort(l1)) << 16);
            }
            if (row_ok) {
              *reinterpret_cast<uint4*>(hrow + 8 * v) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(lrow + 8 * v) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int BN, int PASSES, int EPI>
cudaError_t launch_one(const GemmMaps& m, const GemmParams& p, int num_sms, cudaStream_t st) {
  using C = Cfg<BN, PASSES>;
  auto kern = gemm_tc_kernel<BN, PASSES, EPI>;
  const int n_tiles = ((p.M + BM - 1) / BM) * (p.N / BN);
  const int grid = n_tiles < num_sms ? n_tiles : num_sms;
  kern<<<grid, kThreads, C::kSmemBytes, st>>>(m.a_hi, m.a_lo, m.b_hi, m.b_lo, p);
  return cudaGetLastError();
}

template <int BN, int PASSES, int EPI>
cudaError_t configure_one() {
  return cudaFuncSetAttribute(gemm_tc_kernel<BN, PASSES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              Cfg<BN, PASSES>::kSmemBytes);
}

}  // namespace

// Opt in to >48 KiB dynamic shared memory for every instantiation (once per device, outside graph capture).
cudaError_t configure_gemm_tc() {
  cudaError_t e;
#define D3D_CFG(BN_, PASSES_, EPI_) \
  if ((e = configure_one<BN_, PASSES_, EPI_>()) != cudaSuccess) return e;
  D3D_CFG(128, 3, EPI_F32) D3D_CFG(128, 3, EPI_GELU_SPLIT) D3D_CFG(256, 3, EPI_F32) D3D_CFG(256, 3, EPI_GELU_SPLIT)
  D3D_CFG(128, 1, EPI_F32) D3D_CFG(128, 1, EPI_GELU_SPLIT) D3D_CFG(256, 1, EPI_F32) D3D_CFG(256, 1, EPI_GELU_SPLIT)
#undef D3D_CFG
  return cudaSuccess;
}

cudaError_t launch_gemm_tc(const GemmMaps& maps, const GemmParams& p, int epi, int passes, int bn, int num_sms,
                           cudaStream_t st) {
  if (p.M <= 0) return cudaSuccess;
  if (p.K % BK != 0 || p.N % bn != 0 || (bn != 128 && bn != 256)) return cudaErrorInvalidValue;
#define D3D_DISPATCH(BN_, PASSES_, EPI_) \
  if (bn == BN_ && passes == PASSES_ && epi == EPI_) return launch_one<BN_, PASSES_, EPI_>(maps, p, num_sms, st);
  D3D_DISPATCH(128, 3, EPI_F32)
  D3D_DISPATCH(128, 3, EPI_GELU_SPLIT)
  D3D_DISPATCH(256, 3, EPI_F32)
  D3D_DISPATCH(256, 3, EPI_GELU_SPLIT)
  D3D_DISPATCH(128, 1, EPI_F32)
  D3D_DISPATCH(128, 1, EPI_GELU_SPLIT)
  D3D_DISPATCH(256, 1, EPI_F32)
  D3D_DISPATCH(256, 1, EPI_GELU_SPLIT)
#undef D3D_DISPATCH
  return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------------
// Tensor-map construction (driver entry point fetched through the runtime: no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_operand_map(CUtensorMap* out, const __half* base, int64_t rows, int64_t K) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return -1;
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
  cuuint32_t box[2] = {BK, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

}  // namespace d3d
