// metrics.cu — image metrics of the evaluation loop on the device: PSNR and Gaussian-window SSIM of a rendered frame
// against its target (eval.py:49-84 -> utils/metrics.py:190-197 `eval_errors`: calc_psnr :182-188, ssim(window 11,
// sigma 1.5, zero padding, reduction='mean') :44-126).  The reference builds five full-size filtered maps with grouped
// conv2d calls; here one kernel per frame does the separable 11x11 window in shared memory (16x16 output tiles with a
// 5-pixel halo), writes one (squared error, ssim) partial per block, and a second one-block kernel reduces the
// partials in a fixed order — the result is bit-reproducible run to run.
#include "kernels.h"
#include "profile.h"

namespace mipnerf {
namespace {

constexpr int kTile = 16, kHalo = 5, kWin = 11, kIn = kTile + 2 * kHalo;  // 26

struct GaussWindow {
  float g[kWin];
};

// pred / target: [H, W, C] row-major fp32.  grid = (ceil(W/16), ceil(H/16), C), block = 256.
__global__ void __launch_bounds__(256) ssim_tile_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        int height, int width, int channels, GaussWindow win, float c1,
                                                        float c2, double* __restrict__ partials) {
  __shared__ float sa[kIn][kIn + 1], sb[kIn][kIn + 1];
  __shared__ float hz[5][kIn][kTile + 1];
  __shared__ double red[2][8];
  const int c = blockIdx.z, x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const int tid = threadIdx.x;
  for (int i = tid; i < kIn * kIn; i += 256) {
    const int ly = i / kIn, lx = i % kIn, gy = y0 + ly - kHalo, gx = x0 + lx - kHalo;
    float a = 0.f, b = 0.f;  // zero padding (F.conv2d(padding=5), utils/metrics.py:77)
    if (gy >= 0 && gy < height && gx >= 0 && gx < width) {
      const size_t idx = ((size_t)gy * width + gx) * channels + c;
      a = __ldg(pred + idx);
      b = __ldg(target + idx);
    }
    sa[ly][lx] = a;
    sb[ly][lx] = b;
  }
  __syncthreads();
  for (int i = tid; i < kIn * kTile; i += 256) {  // horizontal pass: 5 maps
    const int ly = i / kTile, lx = i % kTile;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float a = sa[ly][lx + k], b = sb[ly][lx + k], g = win.g[k];
      m1 = fmaf(g, a, m1), m2 = fmaf(g, b, m2);
      s11 = fmaf(g, a * a, s11), s22 = fmaf(g, b * b, s22), s12 = fmaf(g, a * b, s12);
    }
    hz[0][ly][lx] = m1, hz[1][ly][lx] = m2, hz[2][ly][lx] = s11, hz[3][ly][lx] = s22, hz[4][ly][lx] = s12;
  }
  __syncthreads();
  const int ty = tid / kTile, tx = tid % kTile;
  double sq = 0.0, ss = 0.0;
  if (y0 + ty < height && x0 + tx < width) {
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float g = win.g[k];
      m1 = fmaf(g, hz[0][ty + k][tx], m1), m2 = fmaf(g, hz[1][ty + k][tx], m2);
      s11 = fmaf(g, hz[2][ty + k][tx], s11), s22 = fmaf(g, hz[3][ty + k][tx], s22), s12 = fmaf(g, hz[4][ty + k][tx], s12);
    }
    const float mu1_sq = m1 * m1, mu2_sq = m2 * m2, mu12 = m1 * m2;
    const float sig1 = s11 - mu1_sq, sig2 = s22 - mu2_sq, sig12 = s12 - mu12;
    ss = (double)(((2.f * mu12 + c1) * (2.f * sig12 + c2)) / ((mu1_sq + mu2_sq + c1) * (sig1 + sig2 + c2)));
    const float d = sa[ty + kHalo][tx + kHalo] - sb[ty + kHalo][tx + kHalo];
    sq = (double)(d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((tid & 31) == 0) red[0][tid >> 5] = sq, red[1][tid >> 5] = ss;
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) a += red[0][w], b += red[1][w];
    const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    partials[2 * blk] = a;
    partials[2 * blk + 1] = b;
  }
}

// out[0] = psnr (-10 log10 mse), out[1] = mean ssim, out[2] = mse
__global__ void metrics_reduce_kernel(const double* __restrict__ partials, int64_t num_blocks, double count,
                                      float* __restrict__ out) {
  __shared__ double red[2][32];
  double a = 0.0, b = 0.0;
  for (int64_t i = threadIdx.x; i < num_blocks; i += blockDim.x) a += partials[2 * i], b += partials[2 * i + 1];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = a, red[1][threadIdx.x >> 5] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sa += red[0][w], sb += red[1][w];
    const double mse = sa / count;
    out[0] = (float)(-10.0 * log10(mse));
    out[1] = (float)(sb / count);
    out[2] = (float)mse;
  }
}

}  // namespace

size_t image_metrics_scratch_bytes(int height, int width, int channels) {
  const size_t blocks = (size_t)((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile) * channels;
  return blocks * 2 * sizeof(double);
}

cudaError_t launch_image_metrics(const float* pred, const float* target, int height, int width, int channels,
                                 int window, float sigma, float max_val, void* scratch, float* out, cudaStream_t st) {
  if (window != kWin) return cudaErrorInvalidValue;
  GaussWindow win;
  double sum = 0.0;
  float raw[kWin];
  for (int i = 0; i < kWin; ++i) {  // utils/metrics.py:10-17: exp(-(x - 5)^2 / (2 sigma^2)) in fp32, normalised
    const float x = (float)(-(double)((i - kWin / 2) * (i - kWin / 2)) / (2.0 * (double)sigma * (double)sigma));
    raw[i] = expf(x);
    sum += raw[i];
  }
  for (int i = 0; i < kWin; ++i) win.g[i] = raw[i] / (float)sum;
  const float c1 = (0.01f * max_val) * (0.01f * max_val), c2 = (0.03f * max_val) * (0.03f * max_val);
  const dim3 grid((width + kTile - 1) / kTile, (height + kTile - 1) / kTile, channels);
  const int64_t blocks = (int64_t)grid.x * grid.y * grid.z;
  LaunchScope scope(kKernImageMetrics, st);
  ssim_tile_kernel<<<grid, 256, 0, st>>>(pred, target, height, width, channels, win, c1, c2,
                                         static_cast<double*>(scratch));
  metrics_reduce_kernel<<<1, 1024, 0, st>>>(static_cast<const double*>(scratch), blocks,
                                            (double)height * width * channels, out);
  return cudaGetLastError();
}

}  // namespace mipnerf
