// linear_f32.cu — fp32 CUDA-core torch.nn.Linear for the parity path (MIPNERF_B200_FP32).
//
// Y = act([X1 | X2[row / div]] @ W^T + b): the optional second operand expresses both of the
// reference's concatenations without materialising them —
//   * skip connection  cat([h, inputs])            (models/mip_nerf.py:96-97)  div = 1
//   * view condition   cat([bottleneck, viewenc])  (models/mip_nerf.py:106-107) div = samples/ray
// Operands stay fp32 and products are accumulated with FFMA in k order, so the result is within
// fp32 round-off of torch's sgemm; this path exists to demonstrate the 1e-4 parity bar, the
// tensor-core path (mlp_tc.cu) is the fast one.
#include "kernels.h"
#include "profile.h"

namespace mipnerf {

template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x1, int ld1, int k1, const float* __restrict__ x2, int ld2,
                  int k2, int x2_row_div, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ y, int ldy, int64_t m, int n, int relu) {
  static_assert(BM == 128 && BN == 128, "micro-tile mapping assumes 128x128");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  const int K = k1 + k2;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int kk = 0; kk < K; kk += BK) {
#pragma unroll
    for (int i = 0; i < BM * BK / 256; ++i) {
      const int idx = tid + i * 256;
      const int r = idx / BK, k = idx % BK;
      const int64_t row = row0 + r;
      const int kg = kk + k;
      float v = 0.f;
      if (row < m && kg < K)
        v = kg < k1 ? __ldg(x1 + row * ld1 + kg) : __ldg(x2 + (row / x2_row_div) * ld2 + (kg - k1));
      As[k][r] = v;
    }
#pragma unroll
    for (int i = 0; i < BN * BK / 256; ++i) {
      const int idx = tid + i * 256;
      const int r = idx / BK, k = idx % BK;
      const int col = col0 + r;
      const int kg = kk + k;
      Bs[k][r] = (col < n && kg < K) ? __ldg(w + (int64_t)col * K + kg) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = col0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (col >= n) continue;
      float v = acc[i][j] + __ldg(bias + col);
      if (relu) v = fmaxf(v, 0.f);
      y[row * ldy + col] = v;
    }
  }
}

// n <= 8 outputs (density head n=1, colour head n=3): warp per row, lanes stride over k.
__global__ void linear_f32_small_n_kernel(const float* __restrict__ x1, int ld1, int k1,
                                          const float* __restrict__ x2, int ld2, int k2, int x2_row_div,
                                          const float* __restrict__ w, const float* __restrict__ bias,
                                          float* __restrict__ y, int ldy, int64_t m, int n, int relu) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= m) return;
  const int K = k1 + k2;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float xv = k < k1 ? __ldg(x1 + row * ld1 + k) : __ldg(x2 + (row / x2_row_div) * ld2 + (k - k1));
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < n) acc[j] = fmaf(xv, __ldg(w + (int64_t)j * K + k), acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j >= n) break;
    float v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
      v += __ldg(bias + j);
      if (relu) v = fmaxf(v, 0.f);
      y[row * ldy + j] = v;
    }
  }
}

cudaError_t launch_linear_f32(const float* x1, int ld1, int k1, const float* x2, int ld2, int k2,
                              int x2_row_div, const float* w, const float* bias, float* y, int ldy,
                              int64_t m, int n, int relu, cudaStream_t st) {
  if (m == 0 || n == 0) return cudaSuccess;
  if (x2 == nullptr) {
    k2 = 0;
    x2 = x1;
    ld2 = ld1;
  }
  if (x2_row_div < 1) x2_row_div = 1;
  LaunchScope scope(kKernLinearF32, st);
  if (n <= 8) {
    const int warps = 8;
    linear_f32_small_n_kernel<<<(unsigned)((m + warps - 1) / warps), warps * 32, 0, st>>>(
        x1, ld1, k1, x2, ld2, k2, x2_row_div, w, bias, y, ldy, m, n, relu);
  } else {
    dim3 grid((unsigned)((m + 127) / 128), (unsigned)((n + 127) / 128));
    linear_f32_kernel<128, 128, 16><<<grid, 256, 0, st>>>(x1, ld1, k1, x2, ld2, k2, x2_row_div, w, bias, y,
                                                          ldy, m, n, relu);
  }
  return cudaGetLastError();
}

}  // namespace mipnerf
