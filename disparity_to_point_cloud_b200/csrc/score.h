// score.h -- host-side interface of score.cu (matching-score preprocessing of depth_map_fusion)
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace d2pc {

struct ScoreLaunch {
  const uint8_t *frame = nullptr;  // device, mono8 score frame as received (w x h, `step` bytes per row)
  size_t step = 0;
  int width = 0, height = 0;
  bool rotated = false;  // true for score 2: the chain runs on rotateMat(frame) (90 deg clockwise, h cols x w rows)
  int rect[4] = {0, 0, 0, 0};  // cropToSquare rectangle {x, y, n, n} in the (rotated) frame
  uint8_t *out = nullptr;  // device, n x n dense: what the node caches as cropped_score_k_
};
cudaError_t launch_score_preprocess(const ScoreLaunch &L, cudaStream_t stream, int *launches);

}  // namespace d2pc
