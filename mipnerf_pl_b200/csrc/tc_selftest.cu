// tc_selftest.cu — hardware self-test of the tcgen05 building blocks in tc_common.cuh.
//
// D[128,N] = A[128,K] . B[N,K]^T with 16-bit operands and fp32 accumulation in TMEM, through exactly
// the operand layout, descriptors, barriers and TMEM load path the fused kernels use:
//   variant bit 0: B arrives as a pre-swizzled global image via cp.async.bulk + mbarrier tx bytes
//                  (otherwise threads write it with st.shared like the A operand);
//   variant bit 1: A is staged in TMEM with tcgen05.st and the MMA is the TS form.
// tests/test_gpu_tc.py compares D with a torch matmul of the rounded operands.
#include "../../include/mipnerf_b200.h"
#include "profile.h"
#include "tc_common.cuh"

namespace mipnerf {
namespace {

using namespace tc;

template <int kFmt>
__global__ void pack_b_image_kernel(const float* __restrict__ b, uint8_t* __restrict__ image, int n, int k) {
  // image: slabs of [n rows x 128 B], element (row, kk) at sw128_offset(row, kk % 64) in slab kk / 64
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int slabs = (k + 63) / 64;
  if (idx >= n * slabs * 64) return;
  const int row = idx / (slabs * 64);
  const int kk = idx % (slabs * 64);
  const float v = kk < k ? b[(size_t)row * k + kk] : 0.f;
  uint16_t* dst = reinterpret_cast<uint16_t*>(image + (size_t)(kk / 64) * n * 128 + sw128_offset(row, kk % 64));
  *dst = to16<kFmt>(v);
}

template <int kFmt>
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float* __restrict__ a, const float* __restrict__ b, const uint8_t* __restrict__ b_image,
                     float* __restrict__ d, int n, int k, int variant) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);  // SW128 atoms need 1024-B alignment
  const int slabs = (k + 63) / 64;
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)slabs * 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)slabs * n * 128);
  uint64_t* bar_b = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool b_bulk = variant & 1, a_tmem = variant & 2;

  if (tid == 0) {
    mbar_init(bar_b, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);

  // A operand (row = tid) into SW128 slabs, zero padded to the slab width
  for (int s = 0; s < slabs; ++s)
    for (int c = 0; c < 8; ++c) {
      uint32_t w[4];
      for (int q = 0; q < 4; ++q) {
        const int k0 = s * 64 + c * 8 + q * 2;
        const float lo = k0 < k ? a[(size_t)tid * k + k0] : 0.f;
        const float hi = k0 + 1 < k ? a[(size_t)tid * k + k0 + 1] : 0.f;
        w[q] = pack2<kFmt>(lo, hi);
      }
      *reinterpret_cast<uint4*>(sA + (size_t)s * 16384 + sw128_offset(tid, c * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  if (!b_bulk) {
    for (int row = tid; row < n; row += 128)
      for (int s = 0; s < slabs; ++s)
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4];
          for (int q = 0; q < 4; ++q) {
            const int k0 = s * 64 + c * 8 + q * 2;
            const float lo = k0 < k ? b[(size_t)row * k + k0] : 0.f;
            const float hi = k0 + 1 < k ? b[(size_t)row * k + k0 + 1] : 0.f;
            w[q] = pack2<kFmt>(lo, hi);
          }
          *reinterpret_cast<uint4*>(sB + (size_t)s * n * 128 + sw128_offset(row, c * 8)) =
              make_uint4(w[0], w[1], w[2], w[3]);
        }
  }
  fence_proxy_async_smem();  // st.shared operands -> async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (b_bulk && tid == 0) {
    mbar_arrive_expect_tx(bar_b, (uint32_t)(slabs * n * 128));
    for (int s = 0; s < slabs; ++s)
      bulk_g2s(sB + (size_t)s * n * 128, b_image + (size_t)s * n * 128, (uint32_t)(n * 128), bar_b);
  }
  if (a_tmem) {
    // row tid -> TMEM lane tid, columns 256 + kk/2 (two 16-bit elements per 32-bit column)
    for (int c0 = 0; c0 < k / 2; c0 += 16) {
      uint32_t v[16];
      for (int q = 0; q < 16; ++q) {
        const int k0 = (c0 + q) * 2;
        v[q] = pack2<kFmt>(a[(size_t)tid * k + k0], a[(size_t)tid * k + k0 + 1]);
      }
      tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (tid == 0) {
    if (b_bulk) mbar_wait(bar_b, 0);
    const uint32_t idesc = make_idesc_f16(128, n, kFmt);
    uint32_t acc = 0;
    for (int s = 0; s < slabs; ++s) {
      const int steps = ((k - s * 64) < 64 ? (k - s * 64) : 64) / 16;
      for (int j = 0; j < steps; ++j) {
        const uint64_t bdesc = make_sw128_desc(smem_u32(sB + (size_t)s * n * 128) + j * 32);
        if (a_tmem) {
          umma_ts(tmem_base, tmem_base + 256 + (s * 64 + j * 16) / 2, bdesc, idesc, acc);
        } else {
          const uint64_t adesc = make_sw128_desc(smem_u32(sA + (size_t)s * 16384) + j * 32);
          umma_ss(tmem_base, adesc, bdesc, idesc, acc);
        }
        acc = 1;
      }
    }
    umma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  for (int c = 0; c < n; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int q = 0; q < 32; ++q) d[(size_t)tid * n + c + q] = __uint_as_float(v[q]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}


// variant bit 2: one 32-wide K slab in the 64-byte-swizzle layout (A and B via st.shared), k == 32
template <int kFmt>
__global__ void __launch_bounds__(128, 1)
umma_selftest_sw64_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int n) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;             // [128 x 64 B]
  uint8_t* sB = sA + 8192;        // [n x 64 B]
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(sB + (size_t)n * 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  for (int c = 0; c < 4; ++c) {
    uint32_t w[4];
    for (int q = 0; q < 4; ++q) w[q] = pack2<kFmt>(a[tid * 32 + c * 8 + q * 2], a[tid * 32 + c * 8 + q * 2 + 1]);
    *reinterpret_cast<uint4*>(sA + sw64_offset(tid, c * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (int row = tid; row < n; row += 128)
    for (int c = 0; c < 4; ++c) {
      uint32_t w[4];
      for (int q = 0; q < 4; ++q) w[q] = pack2<kFmt>(b[row * 32 + c * 8 + q * 2], b[row * 32 + c * 8 + q * 2 + 1]);
      *reinterpret_cast<uint4*>(sB + sw64_offset(row, c * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, n, kFmt);
    for (int j = 0; j < 2; ++j)
      umma_ss(tmem_base, make_sw64_desc(smem_u32(sA) + j * 32), make_sw64_desc(smem_u32(sB) + j * 32), idesc, j);
    umma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  for (int c = 0; c < n; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int q = 0; q < 32; ++q) d[(size_t)tid * n + c + q] = __uint_as_float(v[q]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// variant bit 3: A as dense K = 16 blocks in the 32-byte-swizzle layout ([k / 16][128 x 32 B]), B as K = 32 slabs in the
// 64-byte-swizzle layout ([k / 32][n x 64 B]) — the operand layouts of the level kernels' activation tile / weight
// stages; k a multiple of 32
template <int kFmt>
__global__ void __launch_bounds__(128, 1)
umma_selftest_a32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int n, int k) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;                          // [k / 16][128 x 32 B]
  uint8_t* sB = sA + (size_t)(k / 16) * 4096;  // [k / 32][n x 64 B]
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(sB + (size_t)(k / 32) * n * 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  for (int c = 0; c < k / 8; ++c) {  // 16-byte chunk c of row tid: block c / 2, half c % 2
    uint32_t w[4];
    for (int q = 0; q < 4; ++q) w[q] = pack2<kFmt>(a[(size_t)tid * k + c * 8 + q * 2], a[(size_t)tid * k + c * 8 + q * 2 + 1]);
    *reinterpret_cast<uint4*>(sA + (size_t)(c >> 1) * 4096 + sw32_offset(tid, (c & 1) * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (int row = tid; row < n; row += 128)
    for (int c = 0; c < k / 8; ++c) {
      uint32_t w[4];
      for (int q = 0; q < 4; ++q) w[q] = pack2<kFmt>(b[(size_t)row * k + c * 8 + q * 2], b[(size_t)row * k + c * 8 + q * 2 + 1]);
      *reinterpret_cast<uint4*>(sB + (size_t)(c >> 2) * n * 64 + sw64_offset(row, (c & 3) * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, n, kFmt);
    for (int j = 0; j < k / 16; ++j)
      umma_ss(tmem_base, make_sw32_desc(smem_u32(sA) + j * 4096),
              make_sw64_desc(smem_u32(sB) + (j >> 1) * n * 64 + (j & 1) * 32), idesc, j ? 1u : 0u);
    umma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  for (int c = 0; c < n; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int q = 0; q < 32; ++q) d[(size_t)tid * n + c + q] = __uint_as_float(v[q]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// Issue-rate microbenchmark: one thread per CTA issues `iters` x 16 back-to-back MMAs (M = 128, N = n, K = 16; a
// 256-deep layer per iteration) on operands that already sit in shared / tensor memory, commits once, and the CTA
// reports clock64 cycles per MMA.  No loads, no epilogue: the ceiling of the MMA pipe for that operand form.
//   mode 0: SS, A = 64-column SW128 slabs (the activation tile of the level kernels), B = SW64 K = 32 stages
//   mode 1: SS, A = dense K = 16 blocks in the 32-byte-swizzle layout, B as above
//   mode 2: TS, A in tensor memory, B as above
template <int kFmt>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int mode, int n, int iters, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;            // 64 KB
  uint8_t* sB = sA + 65536;      // 8 stages x [n x 64 B]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (size_t)8 * n * 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (uint32_t i = tid * 16; i < 65536u + 8u * n * 64u; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, n, kFmt);
    const uint32_t a_u = smem_u32(sA), b_u = smem_u32(sB);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint64_t bd = make_sw64_desc(b_u + s * n * 64 + j * 32);
          if (mode == 2) umma_ts(tmem_base, tmem_base + 256 + (s * 32 + j * 16) / 2, bd, idesc, 1u);
          else if (mode == 1) umma_ss(tmem_base, make_sw32_desc(a_u + (2 * s + j) * 4096), bd, idesc, 1u);
          else umma_ss(tmem_base, make_sw128_desc(a_u + (s >> 1) * 16384 + (s & 1) * 64 + j * 32), bd, idesc, 1u);
        }
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// The same for a CTA pair (cta_group::2, M = 256 over two SMs, each CTA holding n / 2 rows of B): the shapes the level
// kernels issue.  Grid = pairs x 2 with cluster dimension 2; the leader CTA issues and reports.
template <int kFmt>
__global__ void __launch_bounds__(128, 1) umma_rate_pair_kernel(int mode, int n, int iters, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;        // 64 KB
  uint8_t* sB = sA + 65536;  // 8 stages x [n / 2 x 64 B]
  const uint32_t stage = (uint32_t)(n / 2) * 64u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (size_t)8 * stage);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_pair(tmem_slot, 512);
  for (uint32_t i = tid * 16; i < 65536u + 8u * stage; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc_f16(256, n, kFmt);
      const uint32_t a_u = smem_u32(sA), b_u = smem_u32(sB);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t b_lo = desc_lo(b_u + s * stage + j * 32);
            if (mode == 2) umma_ts_pair_lohi(tmem_base, tmem_base + 256 + (s * 32 + j * 16) / 2, b_lo, kDescHiSw64, idesc, 1u);
            else umma_ss_pair_lohi(tmem_base, desc_lo(a_u + (s >> 1) * 16384 + (s & 1) * 64 + j * 32), kDescHiSw128, b_lo,
                                   kDescHiSw64, idesc, 1u);
          }
        }
      }
      umma_commit_pair(bar);
      mbar_wait(bar, 0);
      cycles[blockIdx.x >> 1] = clock64() - t0;
    } else {
      mbar_wait(bar, 0);  // the leader's commit arrives here as well
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace
}  // namespace mipnerf

extern "C" int mipnerf_b200_selftest_umma_rate_pair(int mode, int n, int iters, int precision, int pairs, long long* cycles,
                                                    void* stream) {
  using namespace mipnerf;
  if (!(mode == 0 || mode == 2) || !(n == 128 || n == 256) || iters < 1 || pairs < 1 || !cycles) return MIPNERF_B200_EINVAL;
  if (precision != MIPNERF_B200_BF16 && precision != MIPNERF_B200_FP16) return MIPNERF_B200_EINVAL;
  const size_t sm = 1024 + 65536 + (size_t)8 * (n / 2) * 64 + 64;
  auto kern = precision == MIPNERF_B200_BF16 ? umma_rate_pair_kernel<1> : umma_rate_pair_kernel<0>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return MIPNERF_B200_ECUDA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = sm;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, mode, n, iters, cycles) == cudaSuccess ? MIPNERF_B200_OK : MIPNERF_B200_ECUDA;
}

extern "C" int mipnerf_b200_selftest_umma_rate(int mode, int n, int iters, int precision, int ctas, long long* cycles,
                                               void* stream) {
  using namespace mipnerf;
  if (mode < 0 || mode > 2 || !(n == 128 || n == 256) || iters < 1 || ctas < 1 || !cycles) return MIPNERF_B200_EINVAL;
  if (precision != MIPNERF_B200_BF16 && precision != MIPNERF_B200_FP16) return MIPNERF_B200_EINVAL;
  const size_t sm = 1024 + 65536 + (size_t)8 * n * 64 + 64;
  cudaError_t e;
  if (precision == MIPNERF_B200_BF16) {
    e = cudaFuncSetAttribute(umma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return MIPNERF_B200_ECUDA;
    umma_rate_kernel<1><<<ctas, 128, sm, (cudaStream_t)stream>>>(mode, n, iters, cycles);
  } else {
    e = cudaFuncSetAttribute(umma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return MIPNERF_B200_ECUDA;
    umma_rate_kernel<0><<<ctas, 128, sm, (cudaStream_t)stream>>>(mode, n, iters, cycles);
  }
  return cudaGetLastError() == cudaSuccess ? MIPNERF_B200_OK : MIPNERF_B200_ECUDA;
}

extern "C" int mipnerf_b200_selftest_umma(const float* a, const float* b, float* d, int n, int k, int precision,
                                          int variant, void* scratch, size_t scratch_bytes, void* stream) {
  using namespace mipnerf;
  if (!a || !b || !d || n < 16 || n > 256 || n % 16 || k < 16 || k % 16 || k > 384) return MIPNERF_B200_EINVAL;
  if ((variant & 2) && (k % 32)) return MIPNERF_B200_EINVAL;
  if (precision != MIPNERF_B200_BF16 && precision != MIPNERF_B200_FP16) return MIPNERF_B200_EINVAL;
  if (variant & 8) {
    if (k % 32 || n % 32) return MIPNERF_B200_EINVAL;
    const size_t sm = 1024 + (size_t)(k / 16) * 4096 + (size_t)(k / 32) * n * 64 + 64;
    if (sm > 227 * 1024) return MIPNERF_B200_EUNSUPPORTED;
    cudaError_t e2;
    if (precision == MIPNERF_B200_BF16) {
      e2 = cudaFuncSetAttribute(umma_selftest_a32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      if (e2 != cudaSuccess) return MIPNERF_B200_ECUDA;
      umma_selftest_a32_kernel<1><<<1, 128, sm, (cudaStream_t)stream>>>(a, b, d, n, k);
    } else {
      e2 = cudaFuncSetAttribute(umma_selftest_a32_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      if (e2 != cudaSuccess) return MIPNERF_B200_ECUDA;
      umma_selftest_a32_kernel<0><<<1, 128, sm, (cudaStream_t)stream>>>(a, b, d, n, k);
    }
    return cudaGetLastError() == cudaSuccess ? MIPNERF_B200_OK : MIPNERF_B200_ECUDA;
  }
  if (variant & 4) {
    if (k != 32 || n % 32) return MIPNERF_B200_EINVAL;
    const size_t sm = 1024 + 8192 + (size_t)n * 64 + 64;
    if (precision == MIPNERF_B200_BF16)
      umma_selftest_sw64_kernel<1><<<1, 128, sm, (cudaStream_t)stream>>>(a, b, d, n);
    else
      umma_selftest_sw64_kernel<0><<<1, 128, sm, (cudaStream_t)stream>>>(a, b, d, n);
    return cudaGetLastError() == cudaSuccess ? MIPNERF_B200_OK : MIPNERF_B200_ECUDA;
  }
  const int slabs = (k + 63) / 64;
  const size_t image = (size_t)slabs * n * 128;
  if ((variant & 1) && (!scratch || scratch_bytes < image)) return MIPNERF_B200_EWORKSPACE;
  const size_t smem = 1024 + (size_t)slabs * 16384 + image + 64;
  if (smem > 227 * 1024) return MIPNERF_B200_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const bool bf = precision == MIPNERF_B200_BF16;
  cudaError_t e;
  if (variant & 1) {
    const int total = n * slabs * 64;
    if (bf) pack_b_image_kernel<1><<<(total + 255) / 256, 256, 0, st>>>(b, (uint8_t*)scratch, n, k);
    else pack_b_image_kernel<0><<<(total + 255) / 256, 256, 0, st>>>(b, (uint8_t*)scratch, n, k);
  }
  if (bf) {
    e = cudaFuncSetAttribute(umma_selftest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return MIPNERF_B200_ECUDA;
    umma_selftest_kernel<1><<<1, 128, smem, st>>>(a, b, (const uint8_t*)scratch, d, n, k, variant);
  } else {
    e = cudaFuncSetAttribute(umma_selftest_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return MIPNERF_B200_ECUDA;
    umma_selftest_kernel<0><<<1, 128, smem, st>>>(a, b, (const uint8_t*)scratch, d, n, k, variant);
  }
  return cudaGetLastError() == cudaSuccess ? MIPNERF_B200_OK : MIPNERF_B200_ECUDA;
}
