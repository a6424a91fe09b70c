// linear_tc.cu — a stand-alone tcgen05 linear layer for the training step's forward and dgrad GEMMs:
//   Y[M, N] = epilogue( X[M, K] . B[N, K]^T ),   N in {128, 256},  K in {96, 128, 256},  fp32 in / fp32 out,
// 16-bit operands (bf16 / fp16) rounded while staging, fp32 accumulation in TMEM.
//
// The activations of a training step are HBM-bound (M = rays x samples ~ 5e5 rows, every GEMM reads and writes
// ~0.5 GB) while the weights are tiny, so the kernel is organised around the weights:
//   * persistent CTAs (one per SM); the whole B operand (<= 128 KB, pre-swizzled image written by
//     pack_linear_image_kernel) is fetched ONCE per CTA with cp.async.bulk and stays in shared memory;
//   * per 128-row tile: the 128 threads stage the tile of X cooperatively (coalesced float4 loads, 16 in flight per
//     thread, fp32 -> 16-bit, SW128 K-major slabs: the layout and descriptors validated by mipnerf_b200_selftest_umma), one thread issues K/16 tcgen05.mma (M=128, N, K=16)
//     and commits to an mbarrier, every thread then reads its accumulator row from TMEM and applies the epilogue
//     straight to global memory:  + bias[col]  + row_bias[row / row_div][col]  + prev[row][col]  + r1[row]*r1w[col],
//     ReLU, ReLU-mask by another activation (dgrad);
//   * staging of tile i+1, the MMAs of tile i and the epilogue of tile i-1 overlap (warp-specialised, two TMEM
//     accumulators): the pass is bound by its 1 GB of HBM traffic, not by its 69 GFLOP.
#include <cstdlib>

#include "kernels.h"
#include "profile.h"
#include "tc_common.cuh"

namespace mipnerf {
namespace {

using namespace tc;

// B[n][k] = w[n * ldw + col0 + k]            (transposed == 0: forward, rows of W, columns col0.. of its input dim)
//         = w[(row0 + k) * ldw + n]          (transposed == 1: dgrad, B = W[row0.., :n_dim]^T)
// written as slabs of [n rows x 128 B] in the 128-byte-swizzle K-major layout, zero padded to whole slabs.
template <int kFmt>
__global__ void pack_linear_image_kernel(const float* __restrict__ w, int ldw, int off, int transposed,
                                         uint8_t* __restrict__ image, int n, int k) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int slabs = (k + 63) / 64;
  if (idx >= n * slabs * 64) return;
  const int row = idx / (slabs * 64);
  const int kk = idx % (slabs * 64);
  float v = 0.f;
  if (kk < k) v = transposed ? w[(size_t)(off + kk) * ldw + row] : w[(size_t)row * ldw + off + kk];
  uint16_t* dst = reinterpret_cast<uint16_t*>(image + (size_t)(kk / 64) * n * 128 + sw128_offset(row, kk % 64));
  *dst = to16<kFmt>(v);
}

struct LinearTcParams {
  const float* x;        // [M, ldx], the K columns used start at column 0
  int ldx;
  const uint8_t* image;  // packed B
  float* y;              // [M, ldy]
  int ldy;
  int64_t m;
  int n, k;
  const float* bias;      // [n] or null
  const float* row_bias;  // [M / row_div, n] or null (per-ray view-direction term)
  int row_div;
  const float* prev;      // [M, ldy] or null: added before the activation (second K pass of a concatenated input)
  const float* r1;        // [M] or null, with r1w [n]: rank-1 term r1[row] * r1w[col] (density head in dgrad)
  const float* r1w;
  const float* mask;      // [M, ldy] or null: output zeroed where mask <= 0 (ReLU backward)
  int relu;
};

__device__ __forceinline__ void named_bar(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Warp-specialised pipeline, 256 threads:
//   warps 4-7 ("stagers"): global -> registers -> 16-bit SW128 slabs of tile i+1 while the tensor core works on tile i
//                          and the epilogue warps drain tile i-1; thread 128 also issues the MMAs of the staged tile;
//   warps 0-3 ("epilogue"): thread = accumulator row (TMEM lane quarter = warp), TMEM -> epilogue -> global.
// Two accumulator buffers in TMEM (2 x 256 columns), one A buffer in shared memory:
//   a_free      MMA(i) has consumed the A slabs           (tcgen05.commit)   stagers may overwrite them
//   acc_full[b] MMA(i) has written accumulator b = i & 1  (tcgen05.commit)   epilogue may read it
//   acc_free[b] the epilogue has drained accumulator b    (4 warp arrives)   MMA(i+2) may overwrite it
template <int kFmt>
__global__ void __launch_bounds__(256, 1) linear_tc_kernel(const LinearTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);  // SW128 atoms need 1024-B alignment
  const int slabs = (p.k + 63) / 64;
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)slabs * 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)slabs * p.n * 128);
  uint64_t* bar_b = bars;
  uint64_t* a_free = bars + 1;
  uint64_t* acc_full = bars + 2;  // [2]
  uint64_t* acc_free = bars + 4;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar_b, 1);
    mbar_init(a_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_free[b], 4);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t tiles = (p.m + 127) / 128;

  if (warp >= 4) {
    // ===================================== stagers (+ MMA issue by thread 128) =====================================
    const int st = tid - 128;
    if (st == 0) {  // the weights: once per CTA
      mbar_arrive_expect_tx(bar_b, (uint32_t)(slabs * p.n * 128));
      for (int s = 0; s < slabs; ++s)
        bulk_g2s(sB + (size_t)s * p.n * 128, p.image + (size_t)s * p.n * 128, (uint32_t)(p.n * 128), bar_b);
    }
    const uint32_t idesc = make_idesc_f16(128, p.n, kFmt);
    uint32_t ph_afree = 0, ph_accfree[2] = {0, 0};
    bool b_ready = false;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int64_t row0 = tile * 128;
      for (int s = 0; s < slabs; ++s) {
        float4 f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {  // 16 coalesced loads in flight per thread before anything else
          const int idx = i * 128 + st;
          const int r = idx >> 4, k0 = s * 64 + (idx & 15) * 4;
          f[i] = (row0 + r < p.m && k0 < p.k)
                     ? __ldg(reinterpret_cast<const float4*>(p.x + (row0 + r) * (int64_t)p.ldx + k0))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (s == 0 && it > 0) {  // the previous tile's MMAs must be done with the A slabs
          mbar_wait(a_free, ph_afree);
          ph_afree ^= 1;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int idx = i * 128 + st;
          const int r = idx >> 4, q = idx & 15;
          *reinterpret_cast<uint2*>(sA + (size_t)s * 16384 + sw128_offset(r, (q >> 1) * 8) + (q & 1) * 8) =
              make_uint2(pack2<kFmt>(f[i].x, f[i].y), pack2<kFmt>(f[i].z, f[i].w));
        }
      }
      fence_proxy_async_smem();  // st.shared operand -> async proxy
      named_bar(1, 128);         // all four stager warps have written their part
      if (st == 0) {
        const int buf = it & 1;
        if (!b_ready) {
          mbar_wait(bar_b, 0);
          b_ready = true;
        }
        if (it >= 2) {  // accumulator `buf` was last used by tile it-2: wait until the epilogue drained it
          mbar_wait(&acc_free[buf], ph_accfree[buf]);
          ph_accfree[buf] ^= 1;
        }
        tc_fence_after();
        uint32_t acc = 0;
        for (int s = 0; s < slabs; ++s) {
          const int steps = ((p.k - s * 64) < 64 ? (p.k - s * 64) : 64) / 16;
          for (int j = 0; j < steps; ++j) {
            umma_ss(tmem_base + buf * 256, make_sw128_desc(smem_u32(sA + (size_t)s * 16384) + j * 32),
                    make_sw128_desc(smem_u32(sB + (size_t)s * p.n * 128) + j * 32), idesc, acc);
            acc = 1;
          }
        }
        umma_commit(a_free);
        umma_commit(&acc_full[buf]);
      }
    }
    if (st == 0 && !b_ready) mbar_wait(bar_b, 0);  // never leave with a bulk copy in flight
  } else {
    // ============================================== epilogue warps ==============================================
    uint32_t ph_full[2] = {0, 0};
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int64_t row = tile * 128 + tid;
      const bool valid = row < p.m;
      const float rv = (p.r1 && valid) ? __ldg(p.r1 + row) : 0.f;
      mbar_wait(&acc_full[buf], ph_full[buf]);
      ph_full[buf] ^= 1;
      tc_fence_after();
      for (int c = 0; c < p.n; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 256 + c, v);
        tmem_ld_wait();
        if (valid) {
          float* yr = p.y + row * (int64_t)p.ldy + c;
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            const int col = c + q;
            float o[4] = {__uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]),
                          __uint_as_float(v[q + 3])};
            auto add4 = [&](const float* src) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(src));
              o[0] += t.x, o[1] += t.y, o[2] += t.z, o[3] += t.w;
            };
            if (p.bias) add4(p.bias + col);
            if (p.row_bias) add4(p.row_bias + (row / p.row_div) * p.n + col);
            if (p.prev) {  // may alias y (in-place second K pass): a plain coherent load, not the read-only path
              const float4 t = *reinterpret_cast<const float4*>(p.prev + row * (int64_t)p.ldy + col);
              o[0] += t.x, o[1] += t.y, o[2] += t.z, o[3] += t.w;
            }
            if (p.r1) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(p.r1w + col));
              o[0] = fmaf(rv, t.x, o[0]), o[1] = fmaf(rv, t.y, o[1]), o[2] = fmaf(rv, t.z, o[2]),
              o[3] = fmaf(rv, t.w, o[3]);
            }
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
            }
            if (p.mask) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(p.mask + row * (int64_t)p.ldy + col));
              if (!(t.x > 0.f)) o[0] = 0.f;
              if (!(t.y > 0.f)) o[1] = 0.f;
              if (!(t.z > 0.f)) o[2] = 0.f;
              if (!(t.w > 0.f)) o[3] = 0.f;
            }
            *reinterpret_cast<float4*>(yr + q) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_free[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad on the tensor core:  part[s][n][kg] = sum_{m in slice s} dY[m, n0+n] * Xc[m, kg0+kg],
//   Xc = [X1 (k1 cols) | X2[m / x2_row_div] (k2 cols)], one CTA per (slice, 128-wide n tile, <=256-wide k tile).
// The reduction index is the row m, so BOTH operands are needed "m-major"; they are transposed while staging:
// thread t owns output row n = t of A (and columns t, t+128 of B), reads its column of dY / Xc for the 64 rows of a
// slab (a warp reads 128 contiguous bytes per row: coalesced), and writes eight 16-byte chunks of 8 consecutive m
// into the SW128 K-major slab.  4 MMAs (K = 16) per slab accumulate in TMEM over the whole slice; the bias gradient
// (column sums of dY) falls out of the A staging.  Two CTAs fit per SM (48 KB of shared memory, 256 TMEM columns
// each), so one CTA stages while the other's MMAs run.  Partials have the layout wgrad_reduce_kernel expects.
// ---------------------------------------------------------------------------------------------------------------
struct WgradTcParams {
  const float* dy;  // [M, n_dim]
  int n_dim;
  const float* x1;  // [M, ld1]
  int ld1, k1;
  const float* x2;  // [M / x2_row_div, ld2] or null
  int ld2, k2, x2_row_div;
  float* part;      // [slices, n_dim, K + 1]
  int64_t m, slice_rows;
  // second-generation kernel only: dy / x1 are 16-bit tile images (train_t16.cu) instead of fp32 row-major matrices;
  // x1's / x2's image has ceil(k / 64) slabs per tile (x2 as an image: x2_row_div == 1).
  int dy_t16, x1_t16, x2_t16;
  // optional by-product when x1 is a 256-column tile image: its sign mask, [m][32 bytes], bit c of a row = x1[row][c] > 0.
  // The dgrad GEMM that follows needs exactly this (the ReLU mask of the layer input) and then reads 32 bytes per row
  // instead of the 512-byte activation row again.
  uint8_t* mask_out;
};

template <int kFmt>
__global__ void __launch_bounds__(128, 2) wgrad_tc_kernel(const WgradTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;           // [128 n x 64 m] 16 KB
  uint8_t* sB = sA + 16384;     // [nk   x 64 m] <= 32 KB
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(sB + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = p.k1 + p.k2;
  const int n0 = blockIdx.y * 128, kg0 = blockIdx.z * 256;
  int nk = K - kg0 < 256 ? K - kg0 : 256;  // valid output columns of this k tile
  const int nk_mma = (nk + 15) & ~15;      // UMMA N: a multiple of 16 (padding columns are staged as zeros)
  const int64_t m_begin = (int64_t)blockIdx.x * p.slice_rows;
  const int64_t m_end = (m_begin + p.slice_rows) < p.m ? (m_begin + p.slice_rows) : p.m;

  if (tid == 0) {
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc_f16(128, nk_mma, kFmt);
  uint32_t ph = 0, acc = 0;
  float bsum = 0.f;
  const int n = n0 + tid;
  const bool n_ok = n < p.n_dim;
  auto xcol = [&](int64_t row, int c) -> float {  // Xc[row][c], zero outside
    if (c < p.k1) return __ldg(p.x1 + row * (int64_t)p.ld1 + c);
    if (c < K) return __ldg(p.x2 + (row / p.x2_row_div) * (int64_t)p.ld2 + (c - p.k1));
    return 0.f;
  };
  for (int64_t m0 = m_begin; m0 < m_end; m0 += 64) {
    // ---- A: row n = tid, 64 consecutive m -> 8 chunks of 16 bytes
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int64_t row = m0 + half * 32 + i;
        v[i] = (n_ok && row < m_end) ? __ldg(p.dy + row * (int64_t)p.n_dim + n) : 0.f;
        bsum += v[i];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(sA + sw128_offset(tid, (half * 4 + c) * 8)) =
            make_uint4(pack2<kFmt>(v[c * 8], v[c * 8 + 1]), pack2<kFmt>(v[c * 8 + 2], v[c * 8 + 3]),
                       pack2<kFmt>(v[c * 8 + 4], v[c * 8 + 5]), pack2<kFmt>(v[c * 8 + 6], v[c * 8 + 7]));
    }
    // ---- B: rows (output columns) tid and tid + 128 of this k tile
    for (int rb = tid; rb < nk_mma; rb += 128) {
      const int col = kg0 + rb;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int64_t row = m0 + half * 32 + i;
          v[i] = (rb < nk && row < m_end) ? xcol(row, col) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(sB + sw128_offset(rb, (half * 4 + c) * 8)) =
              make_uint4(pack2<kFmt>(v[c * 8], v[c * 8 + 1]), pack2<kFmt>(v[c * 8 + 2], v[c * 8 + 3]),
                         pack2<kFmt>(v[c * 8 + 4], v[c * 8 + 5]), pack2<kFmt>(v[c * 8 + 6], v[c * 8 + 7]));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        umma_ss(tmem_base, make_sw128_desc(smem_u32(sA) + j * 32), make_sw128_desc(smem_u32(sB) + j * 32), idesc, acc);
        acc = 1;
      }
      umma_commit(bar_mma);
    }
    __syncwarp();
    mbar_wait(bar_mma, ph);  // the slab has been consumed: it may be overwritten
    ph ^= 1;
    tc_fence_after();
  }
  // ---- partial sums of this slice: accumulator row n = tid
  float* out = p.part + (size_t)blockIdx.x * p.n_dim * (K + 1);
  for (int c = 0; c < nk_mma; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    if (n_ok) {
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (c + q < nk) out[(size_t)n * (K + 1) + kg0 + c + q] = m_begin < m_end ? __uint_as_float(v[q]) : 0.f;
    }
  }
  if (blockIdx.z == 0 && n_ok) out[(size_t)n * (K + 1) + K] = bsum;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad, second generation: NO transposition.  The reduction index of dW = dY^T . X is the row m, and a row-major
// [m][col] tile IS the canonical "MN-major" UMMA operand (instruction-descriptor bits 15 / 16): 64 consecutive
// columns (128 B) x 8 rows form one 128-byte-swizzle atom, column blocks LBO apart, 8-row groups SBO apart.  Staging
// is therefore a straight copy: a warp reads one whole row (1 KB, coalesced), rounds to 16 bit and stores one 16-byte
// chunk per lane at the swizzled position of the same row — no per-thread column walks, no 32-row register tiles.
//   grid (slices, k tiles); one CTA per SM covers ALL n_dim <= 256 output rows (two M = 128 accumulators, 2 x 256
//   TMEM columns), so dY and X are each read once per k tile;
//   warps 0-7 stage 64-row slabs into a 3-deep ring (64 KB per stage), warp 8 issues 4 K-steps x n_dim/128 MMAs per
//   slab and frees the stage with tcgen05.commit; after the last slab warps 0-7 drain the accumulators through a
//   shared-memory transpose (coalesced partial rows) and add up the bias gradient (column sums of dY, kept in fp32
//   by the thread that staged the column) in a fixed order.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWgStages = 3;
constexpr int kWgSlab = 64;                    // rows (reduction steps) per stage
constexpr uint32_t kWgOperand = kWgSlab * 512; // 64 rows x 256 cols x 2 B
constexpr uint32_t kWgStage = 2 * kWgOperand;
constexpr uint32_t kWgBlock = kWgSlab * 128;   // one 64-column block of a slab

__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;  // stride between 64-element blocks along M / N
  d |= (uint64_t)(sbo >> 4) << 32;  // stride between 8-row groups along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;           // SWIZZLE_128B
  return d;
}

template <int kFmt>
__device__ __forceinline__ float2 unpack16x2(uint32_t v) {
  if (kFmt == 1) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
  return __half22float2(*reinterpret_cast<__half2*>(&v));
}

template <int kFmt>
__global__ void __launch_bounds__(320, 1) wgrad_mn_kernel(const WgradTcParams p, int swap_strides) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* ring = smem;
  float* bias_s = reinterpret_cast<float*>(ring + kWgStages * kWgStage);  // [8 warps][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 8 * 256);
  uint64_t* full = bars;                  // [stages], 8 arrivals (one per stager warp)
  uint64_t* empty = bars + kWgStages;     // [stages], tcgen05.commit
  uint64_t* done = bars + 2 * kWgStages;  // all MMAs of the slice have completed
  uint64_t* rdone = bars + 2 * kWgStages + 1;  // [stages] tile-image dY: the bias-gradient readers are done with it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kWgStages + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.k1 + p.k2;
  const int kg0 = blockIdx.y * 256;
  const int nk = K - kg0 < 256 ? K - kg0 : 256;  // valid output columns of this k tile
  const int nk_mma = (nk + 63) & ~63;            // whole 64-column blocks (padding staged as zeros)
  const int n_halves = p.n_dim >> 7;
  const int64_t m_begin = (int64_t)blockIdx.x * p.slice_rows;
  const int64_t m_end = (m_begin + p.slice_rows) < p.m ? (m_begin + p.slice_rows) : p.m;
  const bool has_rows = m_begin < m_end;
  const int slabs = has_rows ? (int)((m_end - m_begin + kWgSlab - 1) / kWgSlab) : 0;
  // the k tile lies in X1 or in X2 (the launcher guarantees it never straddles)
  const bool in_x1 = kg0 < p.k1;
  const float* xs = in_x1 ? p.x1 : p.x2;
  const int ldx = in_x1 ? p.ld1 : p.ld2, xdiv = in_x1 ? 1 : p.x2_row_div, xc0 = in_x1 ? kg0 : kg0 - p.k1;
  const bool xvec = (ldx & 3) == 0 && xdiv == 1 && (xc0 & 3) == 0 && (nk & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(xs) & 15) == 0;
  // tile-image operands arrive by bulk copy (warp 9), fp32 operands through the stager warps' registers
  const bool a16 = p.dy_t16 != 0, b16 = in_x1 ? p.x1_t16 != 0 : p.x2_t16 != 0;
  // fp32 operand whose row index is the same for a whole 64-row slab (the per-ray view encodings, one ray = 128 rows):
  // fetched once per slab instead of once per row
  const bool x_per_slab = !b16 && !xvec && (xdiv & 63) == 0;
  const bool bias_read = a16 && blockIdx.y == 0;  // column sums of dY are then taken from the staged tile
  const bool mask_write = p.mask_out != nullptr && b16 && in_x1 && blockIdx.y == 0 && nk == 256;
  const bool reader = bias_read || mask_write;    // the stager warps read staged tile images out of shared memory
  const int a_blocks = p.n_dim >> 6, b_blocks = nk_mma >> 6;

  if (tid == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&full[s], ((a16 && b16) ? 0 : 8) + ((a16 || b16) ? 1 : 0));
      mbar_init(&empty[s], 1);
      mbar_init(&rdone[s], 8);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ================================================ MMA issue ================================================
    if (lane == 0 && has_rows) {
      const uint32_t idesc = make_idesc_f16(128, nk_mma, kFmt) | (1u << 15) | (1u << 16);  // A and B MN-major
      const uint32_t lbo = swap_strides ? 1024u : kWgBlock, sbo = swap_strides ? kWgBlock : 1024u;
      for (int it = 0; it < slabs; ++it) {
        const int s = it % kWgStages;
        mbar_wait(&full[s], (uint32_t)(it / kWgStages) & 1u);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + (size_t)s * kWgStage), sb = sa + kWgOperand;
#pragma unroll
        for (int j = 0; j < kWgSlab / 16; ++j) {  // 16 rows = two 8-row groups = 2 KB further into every block
          for (int h = 0; h < n_halves; ++h)
            umma_ss(tmem_base + h * 256, make_sw128_mn_desc(sa + h * 2 * kWgBlock + j * 2048, lbo, sbo),
                    make_sw128_mn_desc(sb + j * 2048, lbo, sbo), idesc, (it | j) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(done);
    }
  } else if (warp == 9) {
    // ============================== bulk producer (tile-image operands) ==============================
    if (lane == 0 && (a16 || b16)) {
      const uint8_t* dy16 = reinterpret_cast<const uint8_t*>(p.dy);
      const uint8_t* x16 = reinterpret_cast<const uint8_t*>(in_x1 ? p.x1 : p.x2);
      const int x_slabs = ((in_x1 ? p.k1 : p.k2) + 63) >> 6;
      const uint32_t bytes = (uint32_t)((a16 ? a_blocks : 0) + (b16 ? b_blocks : 0)) * kWgBlock;
      // the operand tiles stream through once (0.5 GB per launch): evict-first, so that the 39 MB of partial sums this
      // launch writes are still in L2 when the reduction kernel reads them
      const uint64_t stream_policy = l2_policy_evict_first();
      for (int it = 0; it < slabs; ++it) {
        const int s = it % kWgStages;
        if (it >= kWgStages) {
          const uint32_t par = (uint32_t)(it / kWgStages - 1) & 1u;
          mbar_wait(&empty[s], par);
          if (reader) mbar_wait(&rdone[s], par);
        }
        uint8_t* sa = ring + (size_t)s * kWgStage;
        uint8_t* sb = sa + kWgOperand;
        const int64_t row0 = m_begin + (int64_t)it * kWgSlab;  // a multiple of 64: one half of a 128-row tile
        const size_t tile = (size_t)(row0 >> 7), half = (size_t)((row0 >> 6) & 1) * kWgBlock;
        mbar_arrive_expect_tx(&full[s], bytes);
        if (a16)
          for (int b = 0; b < a_blocks; ++b)
            bulk_g2s_hint(sa + (size_t)b * kWgBlock, dy16 + (tile * a_blocks + b) * 16384 + half, kWgBlock, &full[s],
                          stream_policy);
        if (b16)
          for (int b = 0; b < b_blocks; ++b)
            bulk_g2s_hint(sb + (size_t)b * kWgBlock, x16 + (tile * x_slabs + (xc0 >> 6) + b) * 16384 + half, kWgBlock,
                          &full[s], stream_policy);
      }
    }
  } else {
    // ================================================= stagers =================================================
    float bs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // column sums of dY, columns lane*8 .. +7, this warp's rows
    const int c8 = lane * 8;
    const bool a_on = c8 < p.n_dim, b_on = c8 < nk_mma;
    const uint32_t blk = (uint32_t)(lane >> 3) * kWgBlock, chunk = (uint32_t)(lane & 7);
    for (int it = 0; it < slabs; ++it) {
      const int s = it % kWgStages;
      uint8_t* sa = ring + (size_t)s * kWgStage;
      uint8_t* sb = sa + kWgOperand;
      if (a16 && b16) {
        if (!reader) break;  // nothing for the stager warps to do in this CTA
      } else {
      if (it >= kWgStages) mbar_wait(&empty[s], (uint32_t)(it / kWgStages - 1) & 1u);
      const int64_t m0 = m_begin + (int64_t)it * kWgSlab;
      float4 xs0 = make_float4(0.f, 0.f, 0.f, 0.f), xs1 = xs0;
      if (x_per_slab && c8 < nk) {
        const float* src = xs + (m0 / xdiv) * (int64_t)ldx + xc0 + c8;
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = (c8 + e < nk) ? __ldg(src + e) : 0.f;
        xs0 = make_float4(t[0], t[1], t[2], t[3]), xs1 = make_float4(t[4], t[5], t[6], t[7]);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 fa[4][2], fb[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = warp + 8 * (half * 4 + i);
          const int64_t row = m0 + r;
          const bool ok = row < m_end;
          fa[i][0] = fa[i][1] = fb[i][0] = fb[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && a_on && !a16) {
            const float4* src = reinterpret_cast<const float4*>(p.dy + row * (int64_t)p.n_dim + c8);
            fa[i][0] = __ldg(src), fa[i][1] = __ldg(src + 1);
          }
          if (ok && c8 < nk && !b16) {
            if (x_per_slab) {
              fb[i][0] = xs0, fb[i][1] = xs1;
            } else if (xvec) {
              const float4* src = reinterpret_cast<const float4*>(xs + row * (int64_t)ldx + xc0 + c8);
              fb[i][0] = __ldg(src), fb[i][1] = __ldg(src + 1);
            } else {
              const float* src = xs + (row / xdiv) * (int64_t)ldx + xc0 + c8;
              float t[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) t[e] = (c8 + e < nk) ? __ldg(src + e) : 0.f;
              fb[i][0] = make_float4(t[0], t[1], t[2], t[3]), fb[i][1] = make_float4(t[4], t[5], t[6], t[7]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = warp + 8 * (half * 4 + i);
          const uint32_t off = blk + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4);
          if (a_on && !a16) {
            bs[0] += fa[i][0].x, bs[1] += fa[i][0].y, bs[2] += fa[i][0].z, bs[3] += fa[i][0].w;
            bs[4] += fa[i][1].x, bs[5] += fa[i][1].y, bs[6] += fa[i][1].z, bs[7] += fa[i][1].w;
            *reinterpret_cast<uint4*>(sa + off) =
                make_uint4(pack2<kFmt>(fa[i][0].x, fa[i][0].y), pack2<kFmt>(fa[i][0].z, fa[i][0].w),
                           pack2<kFmt>(fa[i][1].x, fa[i][1].y), pack2<kFmt>(fa[i][1].z, fa[i][1].w));
          }
          if (b_on && !b16)
            *reinterpret_cast<uint4*>(sb + off) =
                make_uint4(pack2<kFmt>(fb[i][0].x, fb[i][0].y), pack2<kFmt>(fb[i][0].z, fb[i][0].w),
                           pack2<kFmt>(fb[i][1].x, fb[i][1].y), pack2<kFmt>(fb[i][1].z, fb[i][1].w));
        }
      }
      fence_proxy_async_smem();  // this thread's st.shared -> visible to the tensor core's (async-proxy) reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      }
      if (reader) {  // operands that arrived as tile images: column sums of dY, sign mask of X, from shared memory
        mbar_wait(&full[s], (uint32_t)(it / kWgStages) & 1u);
        if (mask_write) {
          const int64_t m0r = m_begin + (int64_t)it * kWgSlab;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = warp + 8 * i;
            const uint4 w = *reinterpret_cast<const uint4*>(sb + blk + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4));
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            uint32_t bits = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {  // a 16-bit float is > 0 iff its bits, read as a signed integer, are > 0
              bits |= ((int16_t)(ww[e] & 0xffffu) > 0 ? 1u : 0u) << (2 * e);
              bits |= ((int32_t)ww[e] >= 0x00010000 ? 1u : 0u) << (2 * e + 1);
            }
            p.mask_out[(m0r + r) * 32 + lane] = (uint8_t)bits;  // byte = columns lane*8 .. +7: 32 B per row, coalesced
          }
        }
        if (bias_read && a_on) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = warp + 8 * i;
            const uint4 w = *reinterpret_cast<const uint4*>(sa + blk + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4));
            const float2 f0 = unpack16x2<kFmt>(w.x), f1 = unpack16x2<kFmt>(w.y), f2 = unpack16x2<kFmt>(w.z),
                         f3 = unpack16x2<kFmt>(w.w);
            bs[0] += f0.x, bs[1] += f0.y, bs[2] += f1.x, bs[3] += f1.y;
            bs[4] += f2.x, bs[5] += f2.y, bs[6] += f3.x, bs[7] += f3.y;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&rdone[s]);
      }
    }
    // ---- bias gradient: per-warp column sums -> fixed-order sum over the 8 warps
    if (a_on) {
#pragma unroll
      for (int e = 0; e < 8; ++e) bias_s[warp * 256 + c8 + e] = bs[e];
    }
    if (has_rows) {
      mbar_wait(done, 0);  // every MMA has completed: accumulators final, the ring is free
      tc_fence_after();
    }
    named_bar(1, 256);
    float* out = p.part + (size_t)blockIdx.x * p.n_dim * (K + 1);
    if (blockIdx.y == 0 && tid < p.n_dim) {
      float t = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) t += bias_s[w8 * 256 + tid];
      out[(size_t)tid * (K + 1) + K] = t;
    }
    // ---- accumulators: warp -> (half = warp / 4, TMEM lane quarter = warp % 4), 32 columns at a time through a
    //      padded shared-memory tile so that the global stores run along k (coalesced)
    const int h = warp >> 2, q = warp & 3;
    if (h < n_halves) {
      float* tile = reinterpret_cast<float*>(ring) + warp * (32 * 33);
      const int n_base = h * 128 + q * 32;
      for (int c = 0; c < nk_mma; c += 32) {
        uint32_t v[32];
        if (has_rows) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + h * 256 + c, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0u;
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) tile[lane * 33 + e] = __uint_as_float(v[e]);
        __syncwarp();
        if (c + lane < nk) {
#pragma unroll 8
          for (int r = 0; r < 32; ++r) out[(size_t)(n_base + r) * (K + 1) + kg0 + c + lane] = tile[r * 33 + lane];
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int g_sms = 0;
bool g_attr[2] = {false, false};

}  // namespace

size_t linear_tc_image_bytes(int n, int k) { return (size_t)((k + 63) / 64) * n * 128; }

bool linear_tc_shape_ok(int n, int k) { return (n == 128 || n == 256) && (k == 96 || k == 128 || k == 256); }

cudaError_t launch_pack_linear_image(const float* w, int ldw, int off, int transposed, void* image, int n, int k,
                                     int precision, cudaStream_t st) {
  const int total = n * ((k + 63) / 64) * 64;
  LaunchScope scope(kKernPackWeights, st);
  if (precision == 1)
    pack_linear_image_kernel<1><<<(total + 255) / 256, 256, 0, st>>>(w, ldw, off, transposed, (uint8_t*)image, n, k);
  else
    pack_linear_image_kernel<0><<<(total + 255) / 256, 256, 0, st>>>(w, ldw, off, transposed, (uint8_t*)image, n, k);
  return cudaGetLastError();
}

// precision: 1 = bf16, 2 = fp16 (MIPNERF_B200_BF16 / _FP16)
cudaError_t launch_linear_tc(const float* x, int ldx, const void* image, float* y, int ldy, int64_t m, int n, int k,
                             const float* bias, const float* row_bias, int row_div, const float* prev,
                             const float* r1, const float* r1w, const float* mask, int relu, int precision,
                             cudaStream_t st) {
  if (m == 0) return cudaSuccess;
  if (!linear_tc_shape_ok(n, k) || ldx % 4 != 0 || ldy % 4 != 0) return cudaErrorInvalidValue;
  const int slabs = (k + 63) / 64;
  const size_t smem = 1024 + (size_t)slabs * 16384 + linear_tc_image_bytes(n, k) + 64;
  const int fmt = precision == 1 ? 1 : 0;
  if (!g_attr[fmt]) {
    cudaError_t e = fmt ? cudaFuncSetAttribute(linear_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               200 * 1024)
                        : cudaFuncSetAttribute(linear_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               200 * 1024);
    if (e != cudaSuccess) return e;
    g_attr[fmt] = true;
  }
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  LinearTcParams p{};
  p.x = x, p.ldx = ldx, p.image = static_cast<const uint8_t*>(image), p.y = y, p.ldy = ldy, p.m = m, p.n = n, p.k = k;
  p.bias = bias, p.row_bias = row_bias, p.row_div = row_div < 1 ? 1 : row_div, p.prev = prev;
  p.r1 = r1, p.r1w = r1w, p.mask = mask, p.relu = relu;
  const int64_t tiles = (m + 127) / 128;
  const int grid = (int)(tiles < g_sms ? tiles : g_sms);
  LaunchScope scope(kKernLinearTc, st);
  if (fmt) linear_tc_kernel<1><<<grid, 256, smem, st>>>(p);
  else linear_tc_kernel<0><<<grid, 256, smem, st>>>(p);
  return cudaGetLastError();
}

bool wgrad_tc_shape_ok(int n_dim) { return n_dim == 128 || n_dim == 256; }

// Second-generation wgrad partials.  dy / x1 are fp32 row-major matrices, or (dy_t16 / x1_t16 != 0) 16-bit tile images
// (then m must be a multiple of 128 and x1 has k1 / 64 slabs per tile); x2 is fp32 row-major.
cudaError_t launch_wgrad_mn_partials(const void* dy, int dy_t16, int n_dim, const void* x1, int x1_t16, int ld1, int k1,
                                     const void* x2, int x2_t16, int ld2, int k2, int x2_row_div, float* part,
                                     int64_t m, int max_slices, int precision, int* slices_out, cudaStream_t st,
                                     void* mask_out) {
  if (!x2) x2 = x1, x2_t16 = x1_t16, ld2 = ld1, k2 = 0;
  if (x2_row_div < 1) x2_row_div = 1;
  const int K = k1 + k2;
  if (!(k2 == 0 || k1 % 256 == 0) || !(n_dim == 128 || n_dim == 256)) return cudaErrorInvalidValue;
  if ((dy_t16 || x1_t16 || x2_t16) && m % 128 != 0) return cudaErrorInvalidValue;
  if (x2_t16 && k2 > 0 && x2_row_div != 1) return cudaErrorInvalidValue;
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int fmt = precision == 1 ? 1 : 0;
  const char* swap_env = getenv("MIPNERF_B200_WGRAD_SWAP");
  const int swap_strides = (swap_env && swap_env[0] == '1') ? 1 : 0;
  const int k_tiles = (K + 255) / 256;
  int64_t slices = g_sms / k_tiles;  // one CTA per SM, one wave
  const int64_t by_rows = (m + kWgSlab - 1) / kWgSlab;
  if (slices > by_rows) slices = by_rows;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  int64_t slice_rows = (m + slices - 1) / slices;
  slice_rows = (slice_rows + kWgSlab - 1) / kWgSlab * kWgSlab;
  static bool attr2[2] = {false, false};
  const size_t smem = 1024 + (size_t)kWgStages * kWgStage + 8 * 256 * sizeof(float) + 128;
  if (!attr2[fmt]) {
    cudaError_t e = fmt ? cudaFuncSetAttribute(wgrad_mn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                        : cudaFuncSetAttribute(wgrad_mn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr2[fmt] = true;
  }
  WgradTcParams p{};
  p.dy = static_cast<const float*>(dy), p.n_dim = n_dim, p.x1 = static_cast<const float*>(x1), p.ld1 = ld1, p.k1 = k1;
  p.x2 = static_cast<const float*>(x2), p.ld2 = ld2, p.k2 = k2, p.x2_row_div = x2_row_div, p.part = part, p.m = m;
  p.slice_rows = slice_rows, p.dy_t16 = dy_t16, p.x1_t16 = x1_t16, p.x2_t16 = x2_t16;
  p.mask_out = (x1_t16 && k1 == 256) ? static_cast<uint8_t*>(mask_out) : nullptr;
  dim3 grid((unsigned)slices, (unsigned)k_tiles);
  LaunchScope scope(kKernWgradTc, st);
  if (fmt) wgrad_mn_kernel<1><<<grid, 320, smem, st>>>(p, swap_strides);
  else wgrad_mn_kernel<0><<<grid, 320, smem, st>>>(p, swap_strides);
  *slices_out = (int)slices;
  return cudaGetLastError();
}

// Partials only; the caller runs the fixed-order reduction (train_kernels.cu) afterwards.  Returns the slice count.
// MIPNERF_B200_WGRAD_TC=1 selects the first-generation (transposing) kernel for A/B runs.
cudaError_t launch_wgrad_tc_partials(const float* dy, int n_dim, const float* x1, int ld1, int k1, const float* x2,
                                     int ld2, int k2, int x2_row_div, float* part, int64_t m, int max_slices,
                                     int precision, int* slices_out, cudaStream_t st) {
  if (!x2) x2 = x1, ld2 = ld1, k2 = 0;
  if (x2_row_div < 1) x2_row_div = 1;
  const int K = k1 + k2;
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int fmt = precision == 1 ? 1 : 0;
  const char* gen_env = getenv("MIPNERF_B200_WGRAD_TC");  // read per call: the tests flip it inside one process
  const int generation = (gen_env && gen_env[0] == '1') ? 1 : 2;
  const bool mn_ok = (k2 == 0 || k1 % 256 == 0) && (n_dim & 7) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0;
  if (generation == 2 && mn_ok)
    return launch_wgrad_mn_partials(dy, 0, n_dim, x1, 0, ld1, k1, x2, 0, ld2, k2, x2_row_div, part, m, max_slices,
                                    precision, slices_out, st);
  const int n_tiles = (n_dim + 127) / 128, k_tiles = (K + 255) / 256;
  int64_t slices = (2 * (int64_t)g_sms + n_tiles * k_tiles - 1) / (n_tiles * k_tiles);  // one wave of 2 CTAs per SM
  const int64_t by_rows = (m + 63) / 64;
  if (slices > by_rows) slices = by_rows;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  int64_t slice_rows = (m + slices - 1) / slices;
  slice_rows = (slice_rows + 63) / 64 * 64;
  static bool attr[2] = {false, false};
  const size_t smem = 1024 + 16384 + 32768 + 64;
  if (!attr[fmt]) {
    cudaError_t e = fmt ? cudaFuncSetAttribute(wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                        : cudaFuncSetAttribute(wgrad_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr[fmt] = true;
  }
  WgradTcParams p{};
  p.dy = dy, p.n_dim = n_dim, p.x1 = x1, p.ld1 = ld1, p.k1 = k1, p.x2 = x2, p.ld2 = ld2, p.k2 = k2;
  p.x2_row_div = x2_row_div, p.part = part, p.m = m, p.slice_rows = slice_rows;
  dim3 grid((unsigned)slices, (unsigned)n_tiles, (unsigned)k_tiles);
  LaunchScope scope(kKernWgradTc, st);
  if (fmt) wgrad_tc_kernel<1><<<grid, 128, smem, st>>>(p);
  else wgrad_tc_kernel<0><<<grid, 128, smem, st>>>(p);
  *slices_out = (int)slices;
  return cudaGetLastError();
}

}  // namespace mipnerf
