// Windowing of packed sequences on the device: the F-frame windows that ChunkedGenerator builds on
// the CPU (common/nosiy_generators.py:27-48 for the window bounds, :264-276 for the 2D slice and its flipped copy)
// and the masked write-back of the predictions into per-sequence frame order (target_mask, :264-271, applied in
// RUN:589-596).  Pure gathers / scatters of 8- and 12-byte joints: HBM-bound, one thread per joint.
#include "kernels.cuh"

namespace d3d {
namespace {

// x2d[w, f, j, :] = seq[start[w] + f, j, :];   flip[w, f, j, :] = (-x, y) of seq[start[w] + f, perm[j], :]
__global__ void window_gather_kernel(const float2* __restrict__ seq, const int64_t* __restrict__ start,
                                     const int32_t* __restrict__ perm, float2* __restrict__ x2d,
                                     float2* __restrict__ x2d_flip, int64_t n_joints_total, int F, int J) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_joints_total) return;
  const int j = static_cast<int>(e % J);
  const int64_t wf = e / J;
  const int f = static_cast<int>(wf % F);
  const int64_t w = wf / F;
  const int64_t row = (start[w] + f) * J;
  x2d[e] = seq[row + j];
  if (x2d_flip) {
    float2 v = seq[row + perm[j]];
    v.x = -v.x;
    x2d_flip[e] = v;
  }
}

// seq3d[start[w] + f, j, :] = pred[w, f, j, :]  for f >= first_valid[w]
__global__ void window_scatter_kernel(const float* __restrict__ pred, const int64_t* __restrict__ start,
                                      const int32_t* __restrict__ first_valid, float* __restrict__ seq3d,
                                      int64_t n_joints_total, int F, int J) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_joints_total) return;
  const int j = static_cast<int>(e % J);
  const int64_t wf = e / J;
  const int f = static_cast<int>(wf % F);
  const int64_t w = wf / F;
  if (f < first_valid[w]) return;              // re-predicted overlap of a back-shifted last window
  const int64_t dst = ((start[w] + f) * J + j) * 3;
  seq3d[dst] = pred[e * 3];
  seq3d[dst + 1] = pred[e * 3 + 1];
  seq3d[dst + 2] = pred[e * 3 + 2];
}

}  // namespace

cudaError_t launch_window_gather(const float* seq2d, const int64_t* start, const int32_t* perm, float* x2d,
                                 float* x2d_flip, int64_t n_win, int F, int J, cudaStream_t st) {
  const int64_t n = n_win * F * J;
  if (n <= 0) return cudaSuccess;
  window_gather_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<const float2*>(seq2d), start, perm, reinterpret_cast<float2*>(x2d),
      reinterpret_cast<float2*>(x2d_flip), n, F, J);
  return cudaGetLastError();
}

cudaError_t launch_window_scatter(const float* pred, const int64_t* start, const int32_t* first_valid, float* seq3d,
                                  int64_t n_win, int F, int J, cudaStream_t st) {
  const int64_t n = n_win * F * J;
  if (n <= 0) return cudaSuccess;
  window_scatter_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(pred, start, first_valid, seq3d, n, F, J);
  return cudaGetLastError();
}

}  // namespace d3d
