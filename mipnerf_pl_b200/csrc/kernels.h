// kernels.h — internal (C++) launcher prototypes shared by the translation units of
// libmipnerf_b200.so.  The public surface is include/mipnerf_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mipnerf {

// ---- ray_kernels.cu ----
cudaError_t launch_distloss(const float* weights, const float* t, float* out, int64_t num_rays, int n,
                            cudaStream_t st);
cudaError_t launch_generate_rays(const float* c2w_host, int height, int width, float focal, float near_v,
                                 float far_v, int row0, int rows, float* origins, float* directions,
                                 float* viewdirs, float* radii, float* near_o, float* far_o, cudaStream_t st);
cudaError_t launch_coarse_t(const float* near, const float* far, const float* t_rand, float* t_out,
                            int64_t num_rays, int n, int randomized, int disparity, cudaStream_t st);
cudaError_t launch_cast_rays(const float* origins, const float* directions, const float* radii,
                             const float* t, float* means, float* covs, int64_t num_rays, int n,
                             cudaStream_t st);
cudaError_t launch_ipe(const float* means, const float* covs, float* out, int64_t num_points,
                       int min_deg, int max_deg, cudaStream_t st);
cudaError_t launch_ipe_from_t(const float* origins, const float* directions, const float* radii,
                              const float* t, float* out, int64_t num_rays, int n, int min_deg,
                              int max_deg, int disable_integration, cudaStream_t st);
cudaError_t launch_pos_enc(const float* x, float* out, int64_t num_points, int min_deg, int max_deg,
                           int append_identity, cudaStream_t st);
cudaError_t launch_composite(const float* rgb, const float* dens, const float* t, const float* dirs,
                             float* comp_rgb, float* distance, float* acc, float* weights,
                             int64_t num_rays, int n, int white_bkgd, int activate,
                             float density_bias, float rgb_scale, float rgb_padding, cudaStream_t st);
cudaError_t launch_resample(const float* bins, const float* weights, const float* jitter, float* out,
                            int64_t* inds, int64_t num_rays, int nb, int ns, int randomized, int blur,
                            float padding, cudaStream_t st);

// ---- linear_f32.cu ----
// Y[M,N] = act( [X1 | X2[row / x2_row_div]] @ W[N, K1+K2]^T + bias ),  fp32 FFMA.
cudaError_t launch_linear_f32(const float* x1, int ld1, int k1, const float* x2, int ld2, int k2,
                              int x2_row_div, const float* w, const float* bias, float* y, int ldy,
                              int64_t m, int n, int relu, cudaStream_t st);

}  // namespace mipnerf
