// fusion.h -- host-side interface of fusion.cu (depth_map_fusion's per-pixel merge)
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace d2pc {

// src/depth_map_fusion.cpp:247-265 cropToSquare, for the three rectangles one fusion pass needs.
struct FuseGeometry {
  int r1[4], r2[4], rc[4];  // {x, y, w, h}: map/score 1, map/score 2 (rotated frame), output container
  int n, nc;                // merged square, container square
  int out_x, out_y, out_w, out_h;  // cropMat(left,right,top,bottom) of the container (:130)
};
// returns false where cv::Mat::operator()(Rect) would throw
bool fuse_geometry(int width, int height, int offset_x, int offset_y, int crop_l, int crop_r, int crop_t, int crop_b,
                   FuseGeometry *g);

struct FuseLaunch {
  const uint8_t *d1, *d2, *s1, *s2;  // device, same size / step
  size_t step;
  int width, height;
  FuseGeometry g;
  int rule;
  bool scores_cropped = false;  // s1 / s2 are n x n dense preprocessed scores (merge coordinates)
  uint8_t *container;  // device, nc x nc dense: merge output before the median
  uint8_t *combined;   // device, n x n dense, or nullptr
};
cudaError_t launch_fuse_merge(const FuseLaunch &L, cudaStream_t stream, int *launches);

// DepthMapFusion::colorizeDepth (src/depth_map_fusion.cpp:304-358): gray w x h -> 3 bytes per pixel, dense.
// lut256x4: 1 KB of device memory owned by the caller; build_lut on first use.
cudaError_t launch_colorize(const uint8_t *gray, size_t step, int w, int h, void *lut256x4, bool build_lut,
                            uint8_t *rgb, cudaStream_t stream, int *launches);

}  // namespace d2pc
