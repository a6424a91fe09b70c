// train_t16.cu — the backward pass of the tensor-core training step on 16-bit "tile images".
//
// The training forward (mlp_tc.cu, mlp_level_kernel<.., kTrain>) leaves every activation in HBM exactly as the tensor
// core consumed it: per 128-row tile (= one ray's samples) and 64-column slab a [128 x 128 B] block in the
// 128-byte-swizzle layout (tc::sw128_offset), the slabs of a tile contiguous.  Everything here reads and writes that
// format, so operand staging is one cp.async.bulk per slab — no conversion, no register staging:
//   linear_t16_kernel     dX = [mask > 0] * (dY . B^T + r1[row] * r1w[col])        the dgrad chain, tcgen05
//   color_dgrad_t16       d v = [v > 0] * (d raw_rgb @ Wc)                           colour head -> view layer
//   wgrad_small_n_t16     partials of dY^T X for the two narrow heads (n = 1, 3)     X = tile image, dY fp32
//   t16_pack / t16_unpack fp32 row-major <-> tile image (tests, stand-alone entry points)
// The wgrad GEMMs on tile images are in linear_tc.cu (wgrad_mn_kernel: a row-major tile IS an MN-major operand).
#include "kernels.h"
#include "profile.h"
#include "ray_math.cuh"
#include "tc_common.cuh"

namespace mipnerf {
namespace {

using namespace tc;

constexpr uint32_t kSlab = 16384;  // [128 rows x 64 cols] 16-bit

// 256-bit global accesses (sm_100: LDG.256 / STG.256): one whole 32-byte sector per thread
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

struct LinearT16Params {
  const uint8_t* x;      // [tiles][k / 64][16 KB]
  const uint8_t* image;  // packed B: [k / 64][n x 128 B]   (pack_linear_image_kernel)
  uint8_t* y;            // [tiles][n / 64][16 KB]
  const uint8_t* mask;   // like y, or null: output zeroed where mask <= 0 (ReLU backward)
  const uint8_t* mask_bits;  // or the same mask as sign bits, [tiles * 128][32 B] (n = 256)
  const float* r1;       // [tiles * 128] or null, with r1w [n]: + r1[row] * r1w[col]  (density head)
  const float* r1w;
  int64_t tiles;
  int n, k;
};

// 384 threads: warps 0-3 and 8-11 epilogue (thread = accumulator row, the two groups split the columns), warp 4 lane 0 =
// bulk producer, warp 5 lane 0 = MMA issue.
//   a_full[s] / a_free[s]  per 64-column slab of the A tile: the next tile's slab s streams in as soon as the MMAs of
//                          this tile's slab s have completed (a 4-deep ring at slab granularity, one A buffer);
//   acc_full[b] / acc_free[b]  two TMEM accumulators: the epilogue of tile i overlaps the MMAs of tile i+1.
template <int kFmt>
__global__ void __launch_bounds__(384, 1) linear_t16_kernel(const LinearT16Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int slabs = p.k >> 6;
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)slabs * kSlab;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)slabs * p.n * 128);
  uint64_t* bar_b = bars;
  uint64_t* a_full = bars + 1;     // [4]
  uint64_t* a_free = bars + 5;     // [4]
  uint64_t* acc_full = bars + 9;   // [2]
  uint64_t* acc_free = bars + 11;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar_b, 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_free[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_free[b], 8);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int y_slabs = p.n >> 6;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_b, (uint32_t)(slabs * p.n * 128));
      for (int s = 0; s < slabs; ++s)
        bulk_g2s(sB + (size_t)s * p.n * 128, p.image + (size_t)s * p.n * 128, (uint32_t)(p.n * 128), bar_b);
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        for (int s = 0; s < slabs; ++s) {
          if (it > 0) mbar_wait(&a_free[s], (uint32_t)(it - 1) & 1u);
          mbar_arrive_expect_tx(&a_full[s], kSlab);
          bulk_g2s(sA + (size_t)s * kSlab, p.x + ((size_t)tile * slabs + s) * kSlab, kSlab, &a_full[s]);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, p.n, kFmt);
      mbar_wait(bar_b, 0);
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        if (it >= 2) mbar_wait(&acc_free[buf], (uint32_t)((it >> 1) - 1) & 1u);
        for (int s = 0; s < slabs; ++s) {
          mbar_wait(&a_full[s], (uint32_t)it & 1u);
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            umma_ss(tmem_base + buf * 256, make_sw128_desc(smem_u32(sA + (size_t)s * kSlab) + j * 32),
                    make_sw128_desc(smem_u32(sB + (size_t)s * p.n * 128) + j * 32), idesc, (s | j) ? 1u : 0u);
          umma_commit(&a_free[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp < 4 || warp >= 8) {
    // ============================================== epilogue warps ==============================================
    int it = 0;
    const int row = (warp & 3) * 32 + lane;          // TMEM lane quarter = warp % 4
    const int c_begin = warp < 4 ? 0 : (p.n >> 1), c_end = warp < 4 ? (p.n >> 1) : p.n;
    const uint32_t rx = (uint32_t)row & 7u;
    for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const float rv = p.r1 ? __ldg(p.r1 + tile * 128 + row) : 0.f;
      uint8_t* yrow = p.y + (size_t)tile * y_slabs * kSlab + (uint32_t)row * 128u;
      const uint8_t* mrow = p.mask ? p.mask + (size_t)tile * y_slabs * kSlab + (uint32_t)row * 128u : nullptr;
      uint32_t mb[8] = {~0u, ~0u, ~0u, ~0u, ~0u, ~0u, ~0u, ~0u};
      if (p.mask_bits) ldg256(p.mask_bits + ((size_t)tile * 128 + row) * 32, mb);  // one sector per row per tile
      mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      for (int c = c_begin; c < c_end; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + buf * 256 + c, v);
        // Two adjacent 16-byte chunks of a row share a 32-byte sector of the swizzled tile (positions p and p ^ 1), so
        // the row is read and written one whole sector at a time (256-bit LDG / STG): half the L2 requests of
        // per-chunk accesses.  Which chunk comes first in the sector depends on the row's swizzle bit 0.
        const uint32_t ce0 = (uint32_t)((c & 63) >> 3);  // first (even) chunk of this 32-column group: 0 or 4
        uint32_t mk[2][8];
        if (mrow) {  // issued before the TMEM wait so both latencies overlap
#pragma unroll
          for (int pr = 0; pr < 2; ++pr)
            ldg256(mrow + (size_t)(c >> 6) * kSlab + ((((ce0 + 2 * pr) ^ rx) & ~1u) << 4), mk[pr]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          uint32_t out[8];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {  // chunk ce0 + 2 pr + hf sits in half (hf ^ (rx & 1)) of the sector
            const int j = 2 * pr + hf;
            const uint32_t side = (uint32_t)hf ^ (rx & 1u);
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[8 * j + e]);
            if (p.r1) {
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.r1w + c + 8 * j));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.r1w + c + 8 * j + 4));
              o[0] = fmaf(rv, w0.x, o[0]), o[1] = fmaf(rv, w0.y, o[1]), o[2] = fmaf(rv, w0.z, o[2]);
              o[3] = fmaf(rv, w0.w, o[3]), o[4] = fmaf(rv, w1.x, o[4]), o[5] = fmaf(rv, w1.y, o[5]);
              o[6] = fmaf(rv, w1.z, o[6]), o[7] = fmaf(rv, w1.w, o[7]);
            }
            if (mrow) {  // a 16-bit float is > 0 iff its bits, read as a signed integer, are > 0
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t mw = side ? mk[pr][4 + e] : mk[pr][e];
                if (!((int16_t)(mw & 0xffffu) > 0)) o[2 * e] = 0.f;
                if (!((int32_t)mw >= 0x00010000)) o[2 * e + 1] = 0.f;
              }
            } else if (p.mask_bits) {  // bit (column % 32) of word (column / 32)
              const uint32_t word = mb[c >> 5] >> (8 * j);
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (!((word >> e) & 1u)) o[e] = 0.f;
            }
            const uint32_t w4[4] = {pack2<kFmt>(o[0], o[1]), pack2<kFmt>(o[2], o[3]), pack2<kFmt>(o[4], o[5]),
                                    pack2<kFmt>(o[6], o[7])};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (side) out[4 + e] = w4[e];
              else out[e] = w4[e];
            }
          }
          stg256(yrow + (size_t)(c >> 6) * kSlab + ((((ce0 + 2 * pr) ^ rx) & ~1u) << 4), out);
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

// d v[row][c] = [v[row][c] > 0] * sum_j d_rgb[row][j] * wc[j][c]      thread = (row, 8-column chunk), k_dim = 128
template <int kFmt>
__global__ void color_dgrad_t16_kernel(const float* __restrict__ d_rgb, const float* __restrict__ wc,
                                       const uint8_t* __restrict__ v, uint8_t* __restrict__ d_v, int64_t m,
                                       int k_dim) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunks = k_dim >> 3;
  if (idx >= m * chunks) return;
  const int64_t row = idx / chunks;
  const int ch = (int)(idx % chunks);
  const int r = (int)(row & 127);
  const size_t off = ((size_t)(row >> 7) * (k_dim >> 6) + (ch >> 3)) * kSlab + (uint32_t)r * 128u +
                     ((((uint32_t)ch & 7u) ^ ((uint32_t)r & 7u)) << 4);
  const float g0 = __ldg(d_rgb + row * 3), g1 = __ldg(d_rgb + row * 3 + 1), g2 = __ldg(d_rgb + row * 3 + 2);
  const uint4 mk = __ldg(reinterpret_cast<const uint4*>(v + off));
  const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
  float o[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // Wc rows as float4 (six 16-byte loads instead of 24 scalar ones)
    const int c = ch * 8 + 4 * h;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wc + c));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(wc + k_dim + c));
    const float4 w2 = __ldg(reinterpret_cast<const float4*>(wc + 2 * k_dim + c));
    o[4 * h + 0] = fmaf(g2, w2.x, fmaf(g1, w1.x, g0 * w0.x));
    o[4 * h + 1] = fmaf(g2, w2.y, fmaf(g1, w1.y, g0 * w0.y));
    o[4 * h + 2] = fmaf(g2, w2.z, fmaf(g1, w1.z, g0 * w0.z));
    o[4 * h + 3] = fmaf(g2, w2.w, fmaf(g1, w1.w, g0 * w0.w));
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (!((int16_t)(mw[e] & 0xffffu) > 0)) o[2 * e] = 0.f;
    if (!((int32_t)mw[e] >= 0x00010000)) o[2 * e + 1] = 0.f;
  }
  *reinterpret_cast<uint4*>(d_v + off) =
      make_uint4(pack2<kFmt>(o[0], o[1]), pack2<kFmt>(o[2], o[3]), pack2<kFmt>(o[4], o[5]), pack2<kFmt>(o[6], o[7]));
}

template <int kFmt>
__device__ __forceinline__ float t16_load(const uint8_t* base, int64_t row, int col, int cols) {
  const uint16_t bits = __ldg(reinterpret_cast<const uint16_t*>(
      base + ((size_t)(row >> 7) * (cols >> 6) + (col >> 6)) * kSlab + sw128_offset((int)(row & 127), col & 63)));
  return from16<kFmt>(bits);
}

// The two narrow heads' weight gradients (density n = 1, colour n = 3) with X read from a tile image: an HBM-bound
// streaming pass.  A warp (k_dim = 256) or half-warp (128) owns a row at a time, each lane one 16-byte chunk (8 columns,
// coalesced 512 / 256 B per row), eight rows in flight per lane; the dY values of the row are warp-uniform fp32 loads.
// Lane partials are combined through shared memory in a fixed order; same partial layout as wgrad_small_n_kernel.
template <int kFmt, int kN>
__global__ void __launch_bounds__(256)
wgrad_small_n_t16_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ x, int k_dim,
                         float* __restrict__ part, int64_t m, int64_t slice_rows) {
  __shared__ float red[8][32][kN * 8 + 1];
  __shared__ float bred[16][kN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cpr = k_dim >> 3;        // 16-byte chunks per row: 32 or 16
  const int rpw = 32 / cpr;          // rows a warp covers per step: 1 or 2
  const int sub = lane / cpr, chunk = lane % cpr;
  const int64_t m_begin = (int64_t)blockIdx.x * slice_rows;
  const int64_t m_end = (m_begin + slice_rows) < m ? (m_begin + slice_rows) : m;
  float acc[kN][8], bsum[kN];
#pragma unroll
  for (int j = 0; j < kN; ++j) {
    bsum[j] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
  }
  const size_t slab_off = (size_t)(chunk >> 3) * kSlab;
  const uint32_t ci = (uint32_t)chunk & 7u;
  const int step = 8 * rpw;
  for (int64_t row0 = m_begin + warp * rpw + sub; row0 < m_end; row0 += (int64_t)step * 8) {
    uint4 xv[8];
    float d[8][kN];
#pragma unroll
    for (int u = 0; u < 8; ++u) {  // eight independent rows in flight
      const int64_t row = row0 + (int64_t)u * step;
      const bool ok = row < m_end;
      const int r = (int)(row & 127);
      xv[u] = ok ? __ldg(reinterpret_cast<const uint4*>(x + (size_t)(row >> 7) * (cpr >> 3) * kSlab + slab_off +
                                                        (uint32_t)r * 128u + ((ci ^ ((uint32_t)r & 7u)) << 4)))
                 : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int j = 0; j < kN; ++j) d[u][j] = ok ? __ldg(dy + row * kN + j) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t w[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float x0 = from16<kFmt>((uint16_t)(w[e] & 0xffffu)), x1 = from16<kFmt>((uint16_t)(w[e] >> 16));
#pragma unroll
        for (int j = 0; j < kN; ++j) {
          acc[j][2 * e] = fmaf(d[u][j], x0, acc[j][2 * e]);
          acc[j][2 * e + 1] = fmaf(d[u][j], x1, acc[j][2 * e + 1]);
        }
      }
#pragma unroll
      for (int j = 0; j < kN; ++j) bsum[j] += d[u][j];
    }
  }
#pragma unroll
  for (int j = 0; j < kN; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[warp][lane][j * 8 + e] = acc[j][e];
  if (chunk == 0)
#pragma unroll
    for (int j = 0; j < kN; ++j) bred[warp * 2 + sub][j] = bsum[j];
  __syncthreads();
  float* out = part + (size_t)blockIdx.x * kN * (k_dim + 1);
  if (tid < k_dim) {  // column tid = chunk tid / 8, element tid % 8: fixed-order sum over warps and sub-rows
    const int c = tid >> 3, e = tid & 7;
#pragma unroll
    for (int j = 0; j < kN; ++j) {
      float v = 0.f;
      for (int w8 = 0; w8 < 8; ++w8)
        for (int sb = 0; sb < rpw; ++sb) v += red[w8][sb * cpr + c][j * 8 + e];
      out[(size_t)j * (k_dim + 1) + tid] = v;
    }
  }
  if (tid < kN) {
    float v = 0.f;
    for (int w8 = 0; w8 < 8; ++w8)
      for (int sb = 0; sb < rpw; ++sb) v += bred[w8 * 2 + sb][tid];
    out[(size_t)tid * (k_dim + 1) + k_dim] = v;
  }
}

// The 96 IPE features of every sample as a tile image [rays][2 slabs] (columns 96..127 zero): the X operand of the
// layer-0 / skip-layer wgrads.  Same device functions and the same MUFU fast path as the level kernel's IPE warps
// (mlp_tc.cu: ipe_row_group), so these are bit for bit the features the forward multiplied with.  Thread = (sample row,
// k): k < 6 computes the Gaussian once and the eight (degree, coordinate) pairs 8k..8k+7, i.e. sin chunk k and cos
// chunk 6 + k; k = 6, 7 zero the padding chunks.
template <int kFmt>
__global__ void __launch_bounds__(256) ipe_t16_kernel(const float* __restrict__ origins,
                                                      const float* __restrict__ directions,
                                                      const float* __restrict__ radii, const float* __restrict__ t,
                                                      uint8_t* __restrict__ out, int64_t num_rays, int n,
                                                      int disable_integration) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= num_rays * n * 8) return;
  const int64_t p = idx >> 3;  // sample index = ray * n + j   (n = 128: tile = ray, row = j)
  const int k = (int)(idx & 7);
  const int r = (int)(p & 127);
  uint8_t* tile = out + (size_t)(p >> 7) * (2 * kSlab) + (uint32_t)r * 128u;
  const uint32_t rx = (uint32_t)r & 7u;
  auto chunk_ptr = [&](int ch) { return tile + (size_t)(ch >> 3) * kSlab + ((((uint32_t)ch & 7u) ^ rx) << 4); };
  if (k >= 6) {
    *reinterpret_cast<uint4*>(chunk_ptr(12 + 2 * (k - 6))) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(chunk_ptr(13 + 2 * (k - 6))) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const int64_t ray = p / n;
  const int j = (int)(p % n);
  const RayGeom g = load_ray_geom(origins, directions, radii, ray);
  const float t0 = __ldg(t + ray * (n + 1) + j), t1 = __ldg(t + ray * (n + 1) + j + 1);
  float tm, tv, rv, mean[3], cov[3];
  frustum_moments(t0, t1, g.radius_sq, tm, tv, rv);
  lift_gaussian(g, tm, tv, rv, mean, cov);
  if (disable_integration) cov[0] = cov[1] = cov[2] = 0.f;
  float fs[8], fc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int f = k * 8 + e;  // feature index = degree * 3 + coord   (models/mip.py:335-341)
    ipe_pair<true>(mean[f % 3], cov[f % 3], f / 3, fs[e], fc[e]);
  }
  *reinterpret_cast<uint4*>(chunk_ptr(k)) =
      make_uint4(pack2<kFmt>(fs[0], fs[1]), pack2<kFmt>(fs[2], fs[3]), pack2<kFmt>(fs[4], fs[5]), pack2<kFmt>(fs[6], fs[7]));
  *reinterpret_cast<uint4*>(chunk_ptr(6 + k)) =
      make_uint4(pack2<kFmt>(fc[0], fc[1]), pack2<kFmt>(fc[2], fc[3]), pack2<kFmt>(fc[4], fc[5]), pack2<kFmt>(fc[6], fc[7]));
}

// fp32 row-major [m, cols] (ld) <-> tile image; rows beyond m / columns beyond cols are zero in the image.
// thread = (row, 16-byte chunk of 8 columns): two float4 loads when the source allows it, one 16-byte store
template <int kFmt>
__global__ void t16_pack_kernel(const float* __restrict__ src, int ld, int cols, int64_t m, uint8_t* __restrict__ dst,
                                int img_cols, int64_t padded_rows, int vec) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunks = img_cols >> 3;
  if (idx >= padded_rows * chunks) return;
  const int64_t row = idx / chunks;
  const int ch = (int)(idx % chunks), c0 = ch * 8;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (row < m) {
    if (vec && c0 + 8 <= cols) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + row * ld + c0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + row * ld + c0 + 4));
      v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (c0 + e < cols) v[e] = __ldg(src + row * ld + c0 + e);
    }
  }
  const int r = (int)(row & 127);
  *reinterpret_cast<uint4*>(dst + ((size_t)(row >> 7) * (img_cols >> 6) + (ch >> 3)) * kSlab + (uint32_t)r * 128u +
                            ((((uint32_t)ch & 7u) ^ ((uint32_t)r & 7u)) << 4)) =
      make_uint4(pack2<kFmt>(v[0], v[1]), pack2<kFmt>(v[2], v[3]), pack2<kFmt>(v[4], v[5]), pack2<kFmt>(v[6], v[7]));
}
template <int kFmt>
__global__ void t16_unpack_kernel(const uint8_t* __restrict__ src, int img_cols, float* __restrict__ dst, int ld,
                                  int cols, int64_t m) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * cols) return;
  const int64_t row = idx / cols;
  const int col = (int)(idx % cols);
  dst[row * ld + col] = t16_load<kFmt>(src, row, col, img_cols);
}

int g_sms_t16 = 0;
inline unsigned blocks_of(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }

}  // namespace

size_t t16_image_bytes(int64_t rows, int cols) {
  return (size_t)((rows + 127) / 128) * (size_t)((cols + 63) / 64) * kSlab;
}

cudaError_t launch_t16_pack(const float* src, int ld, int cols, int64_t m, void* image, int precision,
                            cudaStream_t st) {
  const int img_cols = (cols + 63) / 64 * 64;
  const int64_t padded = (m + 127) / 128 * 128;
  if (padded == 0) return cudaSuccess;
  LaunchScope scope(kKernIpe, st);  // accounted with the feature kernels (its use in the training step)
  const int vec = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  const int64_t total = padded * (img_cols / 8);
  if (precision == 1)
    t16_pack_kernel<1><<<blocks_of(total, 256), 256, 0, st>>>(src, ld, cols, m, (uint8_t*)image, img_cols, padded, vec);
  else
    t16_pack_kernel<0><<<blocks_of(total, 256), 256, 0, st>>>(src, ld, cols, m, (uint8_t*)image, img_cols, padded, vec);
  return cudaGetLastError();
}

// min_deg = 0, max_deg = 16 (96 features), n = 128 samples per ray: the level kernels' shape
cudaError_t launch_ipe_t16(const float* origins, const float* directions, const float* radii, const float* t, void* image,
                           int64_t num_rays, int n, int disable_integration, int precision, cudaStream_t st) {
  if (num_rays == 0) return cudaSuccess;
  if (n != 128) return cudaErrorInvalidValue;
  LaunchScope scope(kKernIpe, st);
  const int64_t total = num_rays * n * 8;
  if (precision == 1)
    ipe_t16_kernel<1><<<blocks_of(total, 256), 256, 0, st>>>(origins, directions, radii, t, (uint8_t*)image, num_rays, n,
                                                              disable_integration);
  else
    ipe_t16_kernel<0><<<blocks_of(total, 256), 256, 0, st>>>(origins, directions, radii, t, (uint8_t*)image, num_rays, n,
                                                              disable_integration);
  return cudaGetLastError();
}

cudaError_t launch_t16_unpack(const void* image, int cols, float* dst, int ld, int64_t m, int precision,
                              cudaStream_t st) {
  const int img_cols = (cols + 63) / 64 * 64;
  if (m == 0) return cudaSuccess;
  if (precision == 1)
    t16_unpack_kernel<1><<<blocks_of(m * cols, 256), 256, 0, st>>>((const uint8_t*)image, img_cols, dst, ld, cols, m);
  else
    t16_unpack_kernel<0><<<blocks_of(m * cols, 256), 256, 0, st>>>((const uint8_t*)image, img_cols, dst, ld, cols, m);
  return cudaGetLastError();
}

// y = [mask > 0] * (x . B^T + r1 * r1w) on tile images; m rows (a multiple of 128), n in {128, 256}, k in {128, 256}
cudaError_t launch_linear_t16(const void* x, const void* image, void* y, int64_t m, int n, int k, const float* r1,
                              const float* r1w, const void* mask, int precision, cudaStream_t st,
                              const void* mask_bits) {
  if (m == 0) return cudaSuccess;
  if (m % 128 != 0 || !(n == 128 || n == 256) || !(k == 128 || k == 256)) return cudaErrorInvalidValue;
  if (mask_bits && (mask || n != 256)) return cudaErrorInvalidValue;
  const int slabs = k / 64;
  const size_t smem = 1024 + (size_t)slabs * kSlab + linear_tc_image_bytes(n, k) + 128;
  const int fmt = precision == 1 ? 1 : 0;
  static bool attr[2] = {false, false};
  if (!attr[fmt]) {
    cudaError_t e = fmt ? cudaFuncSetAttribute(linear_t16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                        : cudaFuncSetAttribute(linear_t16_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr[fmt] = true;
  }
  if (g_sms_t16 == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_t16, cudaDevAttrMultiProcessorCount, dev);
  }
  LinearT16Params p{};
  p.x = static_cast<const uint8_t*>(x), p.image = static_cast<const uint8_t*>(image), p.y = static_cast<uint8_t*>(y);
  p.mask_bits = static_cast<const uint8_t*>(mask_bits);
  p.mask = static_cast<const uint8_t*>(mask), p.r1 = r1, p.r1w = r1w, p.tiles = m / 128, p.n = n, p.k = k;
  const int grid = (int)(p.tiles < g_sms_t16 ? p.tiles : g_sms_t16);
  LaunchScope scope(kKernLinearTc, st);
  if (fmt) linear_t16_kernel<1><<<grid, 384, smem, st>>>(p);
  else linear_t16_kernel<0><<<grid, 384, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_color_dgrad_t16(const float* d_rgb, const float* wc, const void* v, void* d_v, int64_t m, int k_dim,
                                   int precision, cudaStream_t st) {
  if (m == 0) return cudaSuccess;
  if (k_dim % 64 != 0) return cudaErrorInvalidValue;
  LaunchScope scope(kKernDgrad, st);
  const int64_t total = m * (k_dim / 8);
  if (precision == 1)
    color_dgrad_t16_kernel<1><<<blocks_of(total, 256), 256, 0, st>>>(d_rgb, wc, (const uint8_t*)v, (uint8_t*)d_v, m, k_dim);
  else
    color_dgrad_t16_kernel<0><<<blocks_of(total, 256), 256, 0, st>>>(d_rgb, wc, (const uint8_t*)v, (uint8_t*)d_v, m, k_dim);
  return cudaGetLastError();
}

// narrow heads: dW[n_dim, k_dim] and db from dy (fp32 [m, n_dim], n_dim = 1 or 3) and a tile-image X; partials +
// fixed-order sum
cudaError_t launch_wgrad_small_n_t16(const float* dy, int n_dim, const void* x, int k_dim, float* part, float* dw,
                                     float* db, int accumulate, int64_t m, int precision, cudaStream_t st,
                                     float scale) {
  if (m == 0 || n_dim == 0) return cudaSuccess;
  if (!(n_dim == 1 || n_dim == 3) || !(k_dim == 128 || k_dim == 256)) return cudaErrorInvalidValue;
  if (g_sms_t16 == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_t16, cudaDevAttrMultiProcessorCount, dev);
  }
  int64_t want = 4 * (int64_t)g_sms_t16;  // four resident blocks per SM, one wave
  if (want > (m + 63) / 64) want = (m + 63) / 64;
  if (want < 1) want = 1;
  const int slices = (int)want;
  const int64_t rows = (m + slices - 1) / slices;
  const uint8_t* x8 = static_cast<const uint8_t*>(x);
  {
    LaunchScope scope(kKernWgrad, st);
    const int fmt = precision == 1 ? 1 : 0;
    if (fmt && n_dim == 1) wgrad_small_n_t16_kernel<1, 1><<<slices, 256, 0, st>>>(dy, x8, k_dim, part, m, rows);
    else if (fmt) wgrad_small_n_t16_kernel<1, 3><<<slices, 256, 0, st>>>(dy, x8, k_dim, part, m, rows);
    else if (n_dim == 1) wgrad_small_n_t16_kernel<0, 1><<<slices, 256, 0, st>>>(dy, x8, k_dim, part, m, rows);
    else wgrad_small_n_t16_kernel<0, 3><<<slices, 256, 0, st>>>(dy, x8, k_dim, part, m, rows);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return launch_wgrad_reduce(part, slices, n_dim, k_dim, dw, db, accumulate, st, scale);
}

}  // namespace mipnerf
