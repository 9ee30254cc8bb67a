// median.h -- host-side launcher interface of median.cu
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "reproject.h"

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
  // Fused DisparityCb (cpp:55-75 in one launch): when `points` is set, each median is scaled (cpp:61), reprojected
  // with the rectified exact arithmetic (reproject_math.cuh) and stored as a PointXYZ at its crop position; `dst` is
  // not written.  The output region must be the crop (ox0 = oy0 = border).  Only for a Q that takes the plain
  // rectified path -- ask reproject_fuses_with_median().
  uint8_t *points = nullptr;  // device: frame f at points + f * points_stride_bytes
  size_t points_stride_bytes = 0;
  const QParams *Q = nullptr;
  float scale = 0.125f;
  bool zero_numer = false;  // Q has an exactly-zero X numerator column: reproject_fuses_with_median() says so
};

cudaError_t launch_median_u8(const MedianLaunch &L, cudaStream_t stream, int *launches);

}  // namespace d2pc
