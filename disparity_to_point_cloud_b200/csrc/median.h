// median.h -- host-side launcher interface of median.cu
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace d2pc {

struct MedianLaunch {
  const uint8_t *src = nullptr;  // device
  uint8_t *dst = nullptr;        // device; addressed with IMAGE coordinates: dst[y*dst_step + x]
  size_t src_step = 0, dst_step = 0, src_frame_stride = 0, dst_frame_stride = 0;
  uint32_t n_frames = 1;
  int width = 0, height = 0;  // image size (the replicate border is the image edge)
  int ox0 = 0, oy0 = 0, ow = 0, oh = 0;  // region of outputs to produce
  int ksize = 11;
  int sm_count = 148;
  int strip_rows = 0;  // 0 = automatic
  int variant = 0;     // 0 = per-thread window histogram (Huang; a selection network for ksize 3),
                       // 2 = window histogram for every ksize (test hook)
};

cudaError_t launch_median_u8(const MedianLaunch &L, cudaStream_t stream, int *launches);

}  // namespace d2pc
