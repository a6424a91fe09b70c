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
//     ReLU, ReLU-mask by another activation (dgrad).
// This is the simple serial pipeline (stage -> MMA -> epilogue per tile): the tensor core idles during staging and
// epilogue, which is acceptable for a pass that is bound by its 1 GB of HBM traffic, not by its 69 GFLOP.
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

template <int kFmt>
__global__ void __launch_bounds__(128, 1) linear_tc_kernel(const LinearTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);  // SW128 atoms need 1024-B alignment
  const int slabs = (p.k + 63) / 64;
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)slabs * 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)slabs * p.n * 128);
  uint64_t* bar_b = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(bar_b, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {  // the weights: once per CTA
    mbar_arrive_expect_tx(bar_b, (uint32_t)(slabs * p.n * 128));
    for (int s = 0; s < slabs; ++s)
      bulk_g2s(sB + (size_t)s * p.n * 128, p.image + (size_t)s * p.n * 128, (uint32_t)(p.n * 128), bar_b);
  }
  const uint32_t idesc = make_idesc_f16(128, p.n, kFmt);
  const int64_t tiles = (p.m + 127) / 128;
  uint32_t ph_mma = 0;
  bool b_ready = false;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row = tile * 128 + tid;
    const bool valid = row < p.m;
    // ---- stage the tile of X, one 64-column slab at a time: 2048 float4 per slab = 16 per thread, consecutive
    //      threads on consecutive 16-byte pieces (coalesced), all 16 loads in flight before the first conversion
    const int64_t row0 = tile * 128;
    for (int s = 0; s < slabs; ++s) {
      float4 f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int idx = i * 128 + tid;
        const int r = idx >> 4, k0 = s * 64 + (idx & 15) * 4;
        f[i] = (row0 + r < p.m && k0 < p.k) ? __ldg(reinterpret_cast<const float4*>(p.x + (row0 + r) * (int64_t)p.ldx + k0))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int idx = i * 128 + tid;
        const int r = idx >> 4, q = idx & 15;
        *reinterpret_cast<uint2*>(sA + (size_t)s * 16384 + sw128_offset(r, (q >> 1) * 8) + (q & 1) * 8) =
            make_uint2(pack2<kFmt>(f[i].x, f[i].y), pack2<kFmt>(f[i].z, f[i].w));
      }
    }
    fence_proxy_async_smem();  // st.shared operand -> async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      if (!b_ready) {
        mbar_wait(bar_b, 0);
        b_ready = true;
      }
      uint32_t acc = 0;
      for (int s = 0; s < slabs; ++s) {
        const int steps = ((p.k - s * 64) < 64 ? (p.k - s * 64) : 64) / 16;
        for (int j = 0; j < steps; ++j) {
          umma_ss(tmem_base, make_sw128_desc(smem_u32(sA + (size_t)s * 16384) + j * 32),
                  make_sw128_desc(smem_u32(sB + (size_t)s * p.n * 128) + j * 32), idesc, acc);
          acc = 1;
        }
      }
      umma_commit(bar_mma);
    }
    __syncwarp();
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    // ---- epilogue: this thread's accumulator row, 32 columns at a time
    const float rv = (p.r1 && valid) ? __ldg(p.r1 + row) : 0.f;
    for (int c = 0; c < p.n; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
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
            o[0] = fmaf(rv, t.x, o[0]), o[1] = fmaf(rv, t.y, o[1]), o[2] = fmaf(rv, t.z, o[2]), o[3] = fmaf(rv, t.w, o[3]);
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
    __syncthreads();  // accumulator drained and A tile consumed: the next tile may overwrite both
    tc_fence_after();
  }
  if (tid == 0 && !b_ready) mbar_wait(bar_b, 0);  // never leave with a bulk copy in flight
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
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
  if (fmt) linear_tc_kernel<1><<<grid, 128, smem, st>>>(p);
  else linear_tc_kernel<0><<<grid, 128, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace mipnerf
