// G-gemm: the qkv / proj / fc1 / fc2 linears of MixSTE (MODEL:75, 84, 51-54) as a persistent,
// warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[M,N] = epilogue( A[M,K] . W[N,K]^T + bias[N] )
//
//   * operands are fp16 "split halves": x = hi + lo with hi = fp16(x), lo = fp16(x - hi).  The default
//     3-pass mode issues   D += A_hi.B_hi ; D += A_hi.B_lo ; D += A_lo.B_hi   per k-step, accumulating in
//     fp32 in TMEM, which restores ~22 mantissa bits (single-pass bf16/fp16 misses the
//     parity bar).  The 1-pass mode issues only A_hi.B_hi.
//   * TMA (cp.async.bulk.tensor, SWIZZLE_128B) stages {128 rows x 64} fp16 boxes; a ring of mbarrier-guarded
//     stages feeds one MMA-issuing thread.
//   * CG = 2 (default): a CTA PAIR (thread-block cluster of 2, tcgen05.mma.cta_group::2) owns a 256 x 256
//     output tile.  Each CTA stages its own 128 rows of A and HALF of the 256 weight rows, so the L2 -> SM
//     operand traffic per MAC is 2/3 of the single-CTA 128 x 256 tile.  ncu on the single-CTA kernel
//     (profiles/r01a) showed lts__throughput at the ~6300 B/clk L2 cap with the tensor pipe only 52 % busy:
//     operand bytes per MAC, not MMA issue, bound this GEMM.  The leader CTA (cluster rank 0) issues the MMAs
//     for both; its tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs' mbarriers.
//     CG = 1 keeps the single-CTA 128 x BN kernel (validation / small problems).
//   * accumulators are double-buffered in TMEM (2 x 256 fp32 columns) so the epilogue of tile i overlaps the
//     mainloop of tile i+1.  Eight epilogue warps per CTA read TMEM with tcgen05.ld (32 lanes x 32 columns);
//     the residual rows of chunk c+1 are prefetched while chunk c is in flight (the un-prefetched version was
//     latency-bound: 26 % tensor-pipe on the proj GEMM).
//
// Warp roles (384 threads): 0 = TMA producer, 1 = MMA issuer (leader CTA), 2 = TMEM allocator, 3 = idle,
// 4..11 = epilogue.
#include "kernels.cuh"
#include "ptx.cuh"

namespace d3d {

namespace {

constexpr int BM = 128;                      // rows per CTA
constexpr int BK = 64;                       // 64 fp16 = 128 B = one swizzle row
constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kTileBytesA = BM * BK * 2;     // 16 KiB
constexpr int kSmemBudget = 200 * 1024;

template <int CG, int BN, int PASSES>
struct Cfg {
  static constexpr int kBRows = BN / CG;                  // weight rows staged by one CTA
  static constexpr int kTileBytesB = kBRows * BK * 2;
  static constexpr int kStageBytes = (PASSES == 3 ? 2 : 1) * (kTileBytesA + kTileBytesB);
  static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (kSmemBudget / kStageBytes);
  static constexpr int kTmemCols = 2 * BN;   // 256 or 512 (power of two)
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(kStages >= 2, "need at least a double buffer");
  static_assert(CG == 1 || BN == 256, "the CTA-pair kernel uses 256 x 256 tiles");
};

struct Barriers {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}

template <int CG, int BN, int PASSES, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
               const GemmParams p) {
  using C = Cfg<CG, BN, PASSES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + C::kStages * C::kStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;      // 0 = leader
  const int unit = CG == 2 ? (blockIdx.x >> 1) : blockIdx.x;        // CTA (pair) index
  const int n_units = CG == 2 ? (gridDim.x >> 1) : gridDim.x;
  constexpr int TM = BM * CG;                                       // tile rows
  const int n_tiles_n = p.N / BN;
  const int n_tiles_m = (p.M + TM - 1) / TM;
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
      ptx::mbar_init(&bars->tmem_empty[a], kEpiWarps * CG);     // leader's copy collects both CTAs' warps
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) ptx::tmem_alloc_cg2<C::kTmemCols>(&bars->tmem_base);
    else ptx::tmem_alloc<C::kTmemCols>(&bars->tmem_base);
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs of a pair)
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit; tile < n_tiles; tile += n_units) {
        const int m0 = (tile / n_tiles_n) * TM + static_cast<int>(rank) * BM;
        const int n0 = (tile % n_tiles_n) * BN + static_cast<int>(rank) * C::kBRows;
        for (int kb = 0; kb < n_kb; ++kb) {
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* s = smem + stage * C::kStageBytes;
          if (CG == 1) {
            ptx::mbar_arrive_expect_tx(&bars->full[stage], C::kStageBytes);
            ptx::tma_load_2d(s, &tm_a_hi, &bars->full[stage], kb * BK, m0);
            s += kTileBytesA;
            if (PASSES == 3) {
              ptx::tma_load_2d(s, &tm_a_lo, &bars->full[stage], kb * BK, m0);
              s += kTileBytesA;
            }
#pragma unroll
            for (int h = 0; h < C::kBRows / 128; ++h)
              ptx::tma_load_2d(s + h * (128 * BK * 2), &tm_b_hi, &bars->full[stage], kb * BK, n0 + h * 128);
            s += C::kTileBytesB;
            if (PASSES == 3) {
#pragma unroll
              for (int h = 0; h < C::kBRows / 128; ++h)
                ptx::tma_load_2d(s + h * (128 * BK * 2), &tm_b_lo, &bars->full[stage], kb * BK, n0 + h * 128);
            }
          } else {
            // both CTAs' bytes complete on the LEADER's full barrier; the leader arms it with the pair's total
            const uint32_t full_leader = ptx::mapa(ptx::smem_u32(&bars->full[stage]), 0);
            if (rank == 0) ptx::mbar_arrive_expect_tx(&bars->full[stage], 2 * C::kStageBytes);
            ptx::tma_load_2d_cg2(s, &tm_a_hi, full_leader, kb * BK, m0);
            s += kTileBytesA;
            if (PASSES == 3) {
              ptx::tma_load_2d_cg2(s, &tm_a_lo, full_leader, kb * BK, m0);
              s += kTileBytesA;
            }
            ptx::tma_load_2d_cg2(s, &tm_b_hi, full_leader, kb * BK, n0);
            s += C::kTileBytesB;
            if (PASSES == 3) ptx::tma_load_2d_cg2(s, &tm_b_lo, full_leader, kb * BK, n0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(TM, BN, 0 /*fp16*/);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = unit; tile < n_tiles; tile += n_units) {
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
            const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
            if (CG == 2) ptx::mma_f16_ss_cg2(d_tmem, da_hi, db_hi, idesc, accum);
            else ptx::mma_f16_ss(d_tmem, da_hi, db_hi, idesc, accum);
            if (PASSES == 3) {
              const uint64_t da_lo = ptx::make_desc_k_sw128(a_lo + k * 32);
              const uint64_t db_lo = ptx::make_desc_k_sw128(b_lo + k * 32);
              if (CG == 2) {
                ptx::mma_f16_ss_cg2(d_tmem, da_hi, db_lo, idesc, 1u);
                ptx::mma_f16_ss_cg2(d_tmem, da_lo, db_hi, idesc, 1u);
              } else {
                ptx::mma_f16_ss(d_tmem, da_hi, db_lo, idesc, 1u);
                ptx::mma_f16_ss(d_tmem, da_lo, db_hi, idesc, 1u);
              }
            }
          }
          // frees the smem stage (in both CTAs) once these MMAs have read it
          if (CG == 2) ptx::mma_commit_cg2(&bars->empty[stage], 3);
          else ptx::mma_commit(&bars->empty[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (CG == 2) ptx::mma_commit_cg2(&bars->tmem_full[acc], 3);
        else ptx::mma_commit(&bars->tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (8 warps per CTA)
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;       // which half of the BN columns
    constexpr int kChunks = BN / 64;        // 32-column chunks per warp
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = unit; tile < n_tiles; tile += n_units) {
      const int m0 = (tile / n_tiles_n) * TM + static_cast<int>(rank) * BM;
      const int n0 = (tile % n_tiles_n) * BN;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const int colbase = n0 + half * (BN / 2);
      const float* rrow = nullptr;
      float4 res[8];
      if (EPI == EPI_F32) {
        if (p.residual && row_ok) rrow = p.residual + static_cast<size_t>(row) * p.N + colbase;
        if (rrow) {                          // prefetch chunk 0 of the residual before the accumulator is ready
#pragma unroll
          for (int v = 0; v < 8; ++v) res[v] = *reinterpret_cast<const float4*>(rrow + 4 * v);
        }
      }
      ptx::mbar_wait(&bars->tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
#pragma unroll
      for (int ci = 0; ci < kChunks; ++ci) {
        const int col0 = half * (BN / 2) + ci * 32;     // column inside the tile
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + acc * BN + col0 + (static_cast<uint32_t>(q * 32) << 16), r);
        float4 nxt[8];
        if (EPI == EPI_F32 && ci + 1 < kChunks && rrow) {
#pragma unroll
          for (int v = 0; v < 8; ++v) nxt[v] = *reinterpret_cast<const float4*>(rrow + (ci + 1) * 32 + 4 * v);
        }
        ptx::tmem_ld_wait();
        const int gcol = n0 + col0;
        const float4* bias4 = reinterpret_cast<const float4*>(p.bias + gcol);
        if (EPI == EPI_F32) {
          float* orow = p.out_f32 + static_cast<size_t>(row) * p.N + gcol;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 b = __ldg(bias4 + v);
            float4 o;
            o.x = __uint_as_float(r[4 * v + 0]) + b.x;
            o.y = __uint_as_float(r[4 * v + 1]) + b.y;
            o.z = __uint_as_float(r[4 * v + 2]) + b.z;
            o.w = __uint_as_float(r[4 * v + 3]) + b.w;
            if (rrow) { o.x += res[v].x; o.y += res[v].y; o.z += res[v].z; o.w += res[v].w; }
            if (row_ok) *reinterpret_cast<float4*>(orow + 4 * v) = o;
          }
          if (ci + 1 < kChunks && rrow) {
#pragma unroll
            for (int v = 0; v < 8; ++v) res[v] = nxt[v];
          }
        } else if (EPI == EPI_GELU_SPLIT) {
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
              hw[e] = pack_h2(h0, h1);
              lw[e] = pack_h2(__float2half_rn(g0 - __half2float(h0)), __float2half_rn(g1 - __half2float(h1)));
            }
            if (row_ok) {
              *reinterpret_cast<uint4*>(hrow + 8 * v) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(lrow + 8 * v) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
        } else {   // EPI_QKV16: q | k -> fp16;  v -> fp16 hi at the same column, lo 512 columns further
          const bool is_v = gcol >= 2 * kC;
          __half* hrow = p.out_qkv + static_cast<size_t>(row) * kQkvRow + gcol;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 b = __ldg(bias4 + 2 * v + (e >> 1));
              const float x0 = __uint_as_float(r[8 * v + 2 * e + 0]) + ((e & 1) ? b.z : b.x);
              const float x1 = __uint_as_float(r[8 * v + 2 * e + 1]) + ((e & 1) ? b.w : b.y);
              const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
              hw[e] = pack_h2(h0, h1);
              lw[e] = pack_h2(__float2half_rn(x0 - __half2float(h0)), __float2half_rn(x1 - __half2float(h1)));
            }
            if (row_ok) {
              *reinterpret_cast<uint4*>(hrow + 8 * v) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              if (is_v) *reinterpret_cast<uint4*>(hrow + kC + 8 * v) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->tmem_empty[acc]), 0));
        else ptx::mbar_arrive(&bars->tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    if (CG == 2) ptx::tmem_dealloc_cg2<C::kTmemCols>(tmem_base);
    else ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int CG, int BN, int PASSES, int EPI>
cudaError_t launch_one(const GemmMaps& m, const GemmParams& p, int num_sms, cudaStream_t st) {
  using C = Cfg<CG, BN, PASSES>;
  auto kern = gemm_tc_kernel<CG, BN, PASSES, EPI>;
  const int n_tiles = ((p.M + BM * CG - 1) / (BM * CG)) * (p.N / BN);
  const int max_units = num_sms / CG;
  const int units = n_tiles < max_units ? n_tiles : max_units;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(units * CG));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, m.a_hi, m.a_lo, m.b_hi, m.b_lo, p);
}

template <int CG, int BN, int PASSES, int EPI>
cudaError_t configure_one() {
  return cudaFuncSetAttribute(gemm_tc_kernel<CG, BN, PASSES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              Cfg<CG, BN, PASSES>::kSmemBytes);
}

}  // namespace

#define D3D_FOR_ALL_GEMMS(X)                                                                        \
  X(1, 128, 3, EPI_F32) X(1, 128, 3, EPI_GELU_SPLIT) X(1, 128, 3, EPI_QKV16)                         \
  X(1, 256, 3, EPI_F32) X(1, 256, 3, EPI_GELU_SPLIT) X(1, 256, 3, EPI_QKV16)                         \
  X(1, 128, 1, EPI_F32) X(1, 128, 1, EPI_GELU_SPLIT) X(1, 128, 1, EPI_QKV16)                         \
  X(1, 256, 1, EPI_F32) X(1, 256, 1, EPI_GELU_SPLIT) X(1, 256, 1, EPI_QKV16)                         \
  X(2, 256, 3, EPI_F32) X(2, 256, 3, EPI_GELU_SPLIT) X(2, 256, 3, EPI_QKV16)                         \
  X(2, 256, 1, EPI_F32) X(2, 256, 1, EPI_GELU_SPLIT) X(2, 256, 1, EPI_QKV16)

// Opt in to >48 KiB dynamic shared memory for every instantiation (once per device, outside graph capture).
cudaError_t configure_gemm_tc() {
  cudaError_t e;
#define D3D_CFG(CG_, BN_, PASSES_, EPI_) \
  if ((e = configure_one<CG_, BN_, PASSES_, EPI_>()) != cudaSuccess) return e;
  D3D_FOR_ALL_GEMMS(D3D_CFG)
#undef D3D_CFG
  return cudaSuccess;
}

cudaError_t launch_gemm_tc(const GemmMaps& maps, const GemmParams& p, int epi, int passes, int bn, int cta_group,
                           int num_sms, cudaStream_t st) {
  if (p.M <= 0) return cudaSuccess;
  if (cta_group == 2) bn = 256;
  if (p.K % BK != 0 || p.N % bn != 0 || (bn != 128 && bn != 256)) return cudaErrorInvalidValue;
  if (epi == EPI_QKV16 && p.N != 3 * kC) return cudaErrorInvalidValue;
#define D3D_DISPATCH(CG_, BN_, PASSES_, EPI_)                                   \
  if (cta_group == CG_ && bn == BN_ && passes == PASSES_ && epi == EPI_)        \
    return launch_one<CG_, BN_, PASSES_, EPI_>(maps, p, num_sms, st);
  D3D_FOR_ALL_GEMMS(D3D_DISPATCH)
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
