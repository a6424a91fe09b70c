// sgemm_tile.cuh — the 128x128x16 fp32 FFMA tile engine shared by the fp32 parity path (linear_f32.cu) and the
// fp32 training step (train_kernels.cu: dgrad, wgrad).
//
// 256 threads, 8x8 micro-tiles (two 4-wide halves per dimension, so shared-memory reads are conflict-free
// float4 broadcasts), one shared-memory stage + register prefetch: the global loads of tile k+1 are issued
// before the 16 FFMA steps of tile k and stored to shared memory after them.  Operand tiles are fetched as
// float4 whenever the caller's strides allow it; every fetch falls back to bounds-checked scalar loads
// (ragged K, the concatenated second operand, the 283-wide view layer).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mipnerf {

constexpr int kTileM = 128, kTileN = 128, kTileK = 16, kTilePad = 4, kTileThreads = 256;

struct TileSmem {
  float a[kTileK][kTileM + kTilePad];  // a[k][row]
  float b[kTileK][kTileN + kTilePad];  // b[k][col]
};

// acc[i][j] += sum_k a[k][row_i] * b[k][col_j];  rows {ty*4..+3, 64+ty*4..+3}, cols likewise with tx
__device__ __forceinline__ void tile_fma(const TileSmem& s, float (&acc)[8][8], int ty, int tx) {
#pragma unroll
  for (int k = 0; k < kTileK; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(&s.a[k][ty * 4]);
    const float4 a1 = *reinterpret_cast<const float4*>(&s.a[k][64 + ty * 4]);
    const float4 b0 = *reinterpret_cast<const float4*>(&s.b[k][tx * 4]);
    const float4 b1 = *reinterpret_cast<const float4*>(&s.b[k][64 + tx * 4]);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

__device__ __forceinline__ int tile_row(int i, int ty) { return i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4); }

// ---- operand fetchers: each thread owns two 4-element groups of a 128x16 (or 16x128) tile -------------------
// "k-contiguous": element (r, k) of the tile lives at src[r * ld + k]; thread group g = tid + 256 i:
//   r = g >> 2, k = (g & 3) * 4 .. +3.  Stored transposed: smem[k][r].
// "row-contiguous": element (k, c) lives at src[k * ld + c]; r/c roles swapped: k = g >> 5, c = (g & 31) * 4 .. +3.
//   Stored as is: smem[k][c..c+3] (one float4).
struct Frag {
  float4 v[2];
};

template <typename F>
__device__ __forceinline__ Frag fetch_frag(F&& f) {  // f(group index 0..511) -> float4
  Frag r;
  r.v[0] = f((int)threadIdx.x);
  r.v[1] = f((int)threadIdx.x + kTileThreads);
  return r;
}

__device__ __forceinline__ void store_kcontig(float (*dst)[kTileM + kTilePad], const Frag& f) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int g = (int)threadIdx.x + i * kTileThreads;
    const int r = g >> 2, k = (g & 3) * 4;
    dst[k + 0][r] = f.v[i].x;
    dst[k + 1][r] = f.v[i].y;
    dst[k + 2][r] = f.v[i].z;
    dst[k + 3][r] = f.v[i].w;
  }
}

__device__ __forceinline__ void store_rowcontig(float (*dst)[kTileN + kTilePad], const Frag& f) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int g = (int)threadIdx.x + i * kTileThreads;
    *reinterpret_cast<float4*>(&dst[g >> 5][(g & 31) * 4]) = f.v[i];
  }
}

// 4 consecutive elements src[0..3] with validity count `n_valid` (elements past it read as 0); `vec` = the caller
// guarantees 16-byte alignment of src when n_valid == 4
__device__ __forceinline__ float4 ld4(const float* __restrict__ src, int n_valid, bool vec) {
  if (vec && n_valid >= 4) return __ldg(reinterpret_cast<const float4*>(src));
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n_valid > 0) r.x = __ldg(src);
  if (n_valid > 1) r.y = __ldg(src + 1);
  if (n_valid > 2) r.z = __ldg(src + 2);
  if (n_valid > 3) r.w = __ldg(src + 3);
  return r;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace mipnerf
