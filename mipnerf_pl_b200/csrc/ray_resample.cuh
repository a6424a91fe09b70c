// ray_resample.cuh — the warp-per-ray inverse-CDF resampler (models/mip.py:168-229, 232-280) as a device
// function, shared by the stand-alone resample kernel (ray_kernels.cu) and the fused level kernel
// (mlp_tc.cu), whose IPE warps resample the fine level's fenceposts in place of a separate launch.
#pragma once
#include <stdint.h>

#include "ray_math.cuh"

namespace mipnerf {

// ---------------------------------------------------------------------------------------------
// inverse-CDF resampling: warp per ray.
// Scratch per warp (shared memory): s_w[nb] (weights -> pdf), s_cdf[nb+1], s_bins[nb+1].
// Bit-exactness notes: the row sum uses torch's 32-strided-accumulator order, the
// CDF is a float64 running sum rounded per prefix; every other op is element-wise IEEE.
// ---------------------------------------------------------------------------------------------
template <bool kBlur>
__device__ __forceinline__ void resample_warp(const float* __restrict__ bins_g,
                                              const float* __restrict__ w_g, int nb, int ns,
                                              int randomized, const Draws& jitter, int64_t ray,
                                              float padding, float* s_w, float* s_cdf,
                                              float* s_bins, float* __restrict__ out_g,
                                              int64_t* __restrict__ inds_g, int lane) {
  const int chunks = nb >> 5;
  for (int i = lane; i <= nb; i += 32) s_bins[i] = __ldg(bins_g + i);
  for (int i = lane; i < nb; i += 32) s_cdf[i] = __ldg(w_g + i);  // raw weights staged in s_cdf
  __syncwarp();
  // blur-pool + padding (models/mip.py:252-257), 32-strided ownership: lane owns 32*i + lane
  float acc = 0.f;
  for (int i = 0; i < chunks; ++i) {
    const int k = 32 * i + lane;
    float w;
    if (kBlur) {
      const float wl = s_cdf[k > 0 ? k - 1 : 0], wc = s_cdf[k], wr = s_cdf[k < nb - 1 ? k + 1 : nb - 1];
      const float m0 = fmaxf(wl, wc), m1 = fmaxf(wc, wr);
      w = __fadd_rn(__fmul_rn(0.5f, __fadd_rn(m0, m1)), padding);
    } else {
      w = s_cdf[k];
    }
    s_w[k] = w;
    acc = i == 0 ? w : __fadd_rn(acc, w);
  }
  // torch.sum order: 4 ILP partials per 8-lane vector, then the 8 lanes in order (:182)
  const int l8 = lane & 7;
  float part = __shfl_sync(0xffffffffu, acc, l8);
  part = __fadd_rn(part, __shfl_sync(0xffffffffu, acc, l8 + 8));
  part = __fadd_rn(part, __shfl_sync(0xffffffffu, acc, l8 + 16));
  part = __fadd_rn(part, __shfl_sync(0xffffffffu, acc, l8 + 24));
  float wsum = __shfl_sync(0xffffffffu, part, 0);
#pragma unroll
  for (int j = 1; j < 8; ++j) wsum = __fadd_rn(wsum, __shfl_sync(0xffffffffu, part, j));
  // eps padding (:183-185)
  const float pad = fmaxf(0.0f, __fsub_rn(1e-5f, wsum));
  const float pad_each = __fdiv_rn(pad, (float)nb);
  wsum = __fadd_rn(wsum, pad);
  bool tiny = false;
  for (int i = 0; i < chunks; ++i) {
    const int k = 32 * i + lane;
    const float pdf = __fdiv_rn(__fadd_rn(s_w[k], pad_each), wsum);  // (:189)
    s_w[k] = pdf;
    tiny |= (pdf != 0.0f && pdf < 1.862645149230957e-09f && k < nb - 1);  // 2^-29
  }
  __syncwarp();
  // cdf = [0, min(1, cumsum(pdf[:-1])), 1]   (:190-195).  With every non-zero pdf >= 2^-29 all
  // float64 partial sums are exact, so the parallel scan equals torch's sequential one bit for
  // bit; otherwise fall back to the sequential order on lane 0.
  if (__any_sync(0xffffffffu, tiny)) {
    if (lane == 0) {
      double run = 0.0;
      s_cdf[0] = 0.0f;
      for (int k = 0; k < nb - 1; ++k) {
        run += (double)s_w[k];
        s_cdf[k + 1] = fminf(1.0f, (float)run);
      }
      s_cdf[nb] = 1.0f;
    }
  } else {
    double run = 0.0;
    for (int p = 0; p < chunks; ++p) {
      const int k = lane * chunks + p;
      if (k < nb - 1) run += (double)s_w[k];
    }
    double total;
    double before = warp_excl_scan_f64(run, lane, total);
    for (int p = 0; p < chunks; ++p) {
      const int k = lane * chunks + p;
      if (k < nb - 1) {
        before += (double)s_w[k];
        s_cdf[k + 1] = fminf(1.0f, (float)before);
      }
    }
    if (lane == 0) {
      s_cdf[0] = 0.0f;
      s_cdf[nb] = 1.0f;
    }
  }
  __syncwarp();
  // u (:198-208), searchsorted(right=True) (:211), interpolation (:219-228)
  const float one_m_eps = 1.0f - MIPNERF_F32_EPS;
  const float step = (float)(1.0 / (double)ns);  // randomized: arange(ns) * (1/ns)   (models/mip.py:198-200)
  for (int j = lane; j < ns; j += 32) {
    // deterministic: torch.linspace(0, 1-eps, ns)   (models/mip.py:206-207)
    float u = randomized ? __fmul_rn((float)j, step) : linspace_f32(0.0f, one_m_eps, ns, j);
    if (randomized) u = fminf(__fadd_rn(u, draw_uniform(jitter, ray, j, ns)), one_m_eps);
    int lo = 0, hi = nb + 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_cdf[mid] <= u) lo = mid + 1;
      else hi = mid;
    }
    const int below = lo - 1 > 0 ? lo - 1 : 0;
    const int above = lo < nb ? lo : nb;
    const float cb = s_cdf[below], ca = s_cdf[above];
    const float bb = s_bins[below], ba = s_bins[above];
    float denom = __fsub_rn(ca, cb);
    if (denom < 1e-5f) denom = 1.0f;
    const float tt = __fdiv_rn(__fsub_rn(u, cb), denom);
    out_g[j] = __fadd_rn(bb, __fmul_rn(tt, __fsub_rn(ba, bb)));
    if (inds_g) inds_g[j] = (int64_t)lo;
  }
}


// Same arithmetic, ONE scratch array of nb+1 floats (bins are re-read from global memory): for callers that are
// short of shared memory (the fused level kernel).  The array holds, in turn, the raw weights (blur-pool
// neighbours), the pdf in 32-strided order, and finally the cdf.  Results are bit-identical to resample_warp.
template <bool kBlur>
__device__ __forceinline__ void resample_warp_lean(const float* __restrict__ bins_g, const float* __restrict__ w_g,
                                                   int nb, int ns, int randomized, const Draws& jitter,
                                                   int64_t ray, float padding, float* s_a,
                                                   float* __restrict__ out_g, int64_t* __restrict__ inds_g, int lane) {
  constexpr int kMaxChunks = 8;  // nb <= 256
  const int chunks = nb >> 5;
  for (int i = lane; i < nb; i += 32) s_a[i] = __ldg(w_g + i);
  __syncwarp();
  float w[kMaxChunks];
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    if (i < chunks) {
      const int k = 32 * i + lane;
      if (kBlur) {
        const float wl = s_a[k > 0 ? k - 1 : 0], wc = s_a[k], wr = s_a[k < nb - 1 ? k + 1 : nb - 1];
        const float m0 = fmaxf(wl, wc), m1 = fmaxf(wc, wr);
        w[i] = __fadd_rn(__fmul_rn(0.5f, __fadd_rn(m0, m1)), padding);
      } else {
        w[i] = s_a[k];
      }
      acc = i == 0 ? w[i] : __fadd_rn(acc, w[i]);
    }
  }
  const int l8 = lane & 7;
  float part = __shfl_sync(0xffffffffu, acc, l8);
  part = __fadd_rn(part, __shfl_sync(0xffffffffu, acc, l8 + 8));
  part = __fadd_rn(part, __shfl_sync(0xffffffffu, acc, l8 + 16));
  part = __fadd_rn(part, __shfl_sync(0xffffffffu, acc, l8 + 24));
  float wsum = __shfl_sync(0xffffffffu, part, 0);
#pragma unroll
  for (int j = 1; j < 8; ++j) wsum = __fadd_rn(wsum, __shfl_sync(0xffffffffu, part, j));
  const float pad = fmaxf(0.0f, __fsub_rn(1e-5f, wsum));
  const float pad_each = __fdiv_rn(pad, (float)nb);
  wsum = __fadd_rn(wsum, pad);
  bool tiny = false;
  __syncwarp();  // every lane is done with the raw weights
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    if (i < chunks) {
      const int k = 32 * i + lane;
      const float pdf = __fdiv_rn(__fadd_rn(w[i], pad_each), wsum);
      s_a[k] = pdf;
      tiny |= (pdf != 0.0f && pdf < 1.862645149230957e-09f && k < nb - 1);  // 2^-29
    }
  }
  __syncwarp();
  if (__any_sync(0xffffffffu, tiny)) {
    if (lane == 0) {  // torch's sequential order, in place: s_a[k] (pdf) -> s_a[k+1] (cdf)
      double run = 0.0;
      float cur = s_a[0];
      s_a[0] = 0.0f;
      for (int k = 0; k < nb - 1; ++k) {
        run += (double)cur;
        cur = s_a[k + 1];
        s_a[k + 1] = fminf(1.0f, (float)run);
      }
      s_a[nb] = 1.0f;
    }
  } else {
    float pdf_b[kMaxChunks];
    double run = 0.0;
#pragma unroll
    for (int q = 0; q < kMaxChunks; ++q) {
      if (q < chunks) {
        const int k = lane * chunks + q;
        pdf_b[q] = s_a[k];
        if (k < nb - 1) run += (double)pdf_b[q];
      }
    }
    double total;
    double before = warp_excl_scan_f64(run, lane, total);
    __syncwarp();  // all pdf values are in registers: the array may now take the cdf
#pragma unroll
    for (int q = 0; q < kMaxChunks; ++q) {
      if (q < chunks) {
        const int k = lane * chunks + q;
        if (k < nb - 1) {
          before += (double)pdf_b[q];
          s_a[k + 1] = fminf(1.0f, (float)before);
        }
      }
    }
    if (lane == 0) {
      s_a[0] = 0.0f;
      s_a[nb] = 1.0f;
    }
  }
  __syncwarp();
  const float one_m_eps = 1.0f - MIPNERF_F32_EPS;
  const float step = (float)(1.0 / (double)ns);  // randomized: arange(ns) * (1/ns)   (models/mip.py:198-200)
  for (int j = lane; j < ns; j += 32) {
    // deterministic: torch.linspace(0, 1-eps, ns)   (models/mip.py:206-207)
    float u = randomized ? __fmul_rn((float)j, step) : linspace_f32(0.0f, one_m_eps, ns, j);
    if (randomized) u = fminf(__fadd_rn(u, draw_uniform(jitter, ray, j, ns)), one_m_eps);
    int lo = 0, hi = nb + 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_a[mid] <= u) lo = mid + 1;
      else hi = mid;
    }
    const int below = lo - 1 > 0 ? lo - 1 : 0;
    const int above = lo < nb ? lo : nb;
    const float cb = s_a[below], ca = s_a[above];
    const float bb = __ldg(bins_g + below), ba = __ldg(bins_g + above);
    float denom = __fsub_rn(ca, cb);
    if (denom < 1e-5f) denom = 1.0f;
    const float tt = __fdiv_rn(__fsub_rn(u, cb), denom);
    out_g[j] = __fadd_rn(bb, __fmul_rn(tt, __fsub_rn(ba, bb)));
    if (inds_g) inds_g[j] = (int64_t)lo;
  }
  __syncwarp();
}

}  // namespace mipnerf
