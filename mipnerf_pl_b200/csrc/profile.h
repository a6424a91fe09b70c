// profile.h — per-kernel launch counters and optional CUDA-event timing inside the library.
// bench.py uses it for `gpu_launches` and for the dominant kernel's live launch duration
// (`roofline.achieved`); with timing disabled the only cost is one counter increment per launch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mipnerf {

enum KernelId : int {
  kKernCoarseT = 0,
  kKernCastRays,
  kKernIpe,
  kKernPosEnc,
  kKernLinearF32,
  kKernComposite,
  kKernResample,
  kKernPackWeights,
  kKernMlpLevelTc,   // fused tcgen05 level kernel (IPE + MLP + compositing)
  kKernMlpTc,        // tcgen05 MLP on explicit features
  kKernRayGen,       // on-device pinhole ray generation
  kKernDistloss,
  kKernRayPrologue,  // view-direction bias + coarse fenceposts in front of the fused level kernels
  kKernRenderBackward,
  kKernDgrad,
  kKernWgrad,
  kKernAdam,
  kKernLinearTc,     // stand-alone tcgen05 linear layer (training forward / dgrad)
  kKernWgradTc,      // tcgen05 wgrad partials
  kKernImageMetrics, // PSNR + SSIM of a rendered frame
  kKernCount
};

const char* kernel_name(int id);

// RAII bracket around one kernel launch on `st`.
struct LaunchScope {
  LaunchScope(int id, cudaStream_t st);
  ~LaunchScope();
  int id_;
  cudaStream_t st_;
  int slot_;
};

}  // namespace mipnerf
