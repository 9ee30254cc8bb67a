// fusion.cu -- depth_map_fusion's per-pixel merge on sm_100a.
//
// Replaces the loop of DepthMapFusion::publishFusedDepthMap
//   src/depth_map_fusion.cpp:113-123   (merge + combined score)
//   :150-160 getFusedDistance, :219-235 gradFilter, :169-217 alternate rules
// with the geometry of :237-273 (cropToSquare / rotateMat / cropMat) folded
// into index arithmetic: no rotated or cropped copy is ever materialised.
//   cropped_1(i,j) = src1(y1+i, x1+j)
//   cropped_2(i,j) = rot(y2+i, x2+j) = src2(H-1-(x2+j), y2+i)      rot(r,c)=src(H-1-c,r)
// Map 2 / score 2 are walked along source rows (coalesced) and transposed
// through shared memory, 32x32 pixels per CTA.
// HBM traffic: 4 bytes read + 2 bytes written per merged pixel (6 n^2).
#include "fusion.h"

#include <algorithm>
#include <cstdlib>

namespace d2pc {
namespace {

struct FuseArgs {
  const uint8_t *d1, *d2, *s1, *s2;
  size_t step;
  int width, height;
  int x1, y1, x2, y2, xc, yc, n, nc;
  int rule;
  int scores_cropped;  // s1 / s2 are n x n dense images already in merge coordinates (preprocessed caches)
  uint8_t *container, *combined;
};

// src/depth_map_fusion.cpp:219-235.  The two double comparisons are restated on the float ratio:
//   0.8 < (double)r   <=>  r >= 0x1.99999ap-1f  (float(0.8) is the first float above 0.8)
//   (double)r < 1.25  <=>  r < 1.25f            (1.25 is exact)
// NaN (0/0) and inf (d/0) fail both, as in the reference.  The division is IEEE (no fast-math).
__device__ __forceinline__ int grad_filter(int d1, int d2, int s1, int s2) {
  if (s1 < s2 && s1 < 100 && d1 < 230) return d1;
  if (s2 < s1 && s2 < 100 && d2 < 230) return d2;
  const float r = __fdiv_rn((float)d1, (float)d2);
  if (r >= 0x1.99999ap-1f && r < 1.25f && s1 < 125 && s2 < 125) return (d1 + d2) >> 1;
  return 0;
}

__device__ __forceinline__ int fuse_rule(int rule, int d1, int d2, int s1, int s2) {
  switch (rule) {
    case 0: return grad_filter(d1, d2, s1, s2);
    case 1: return min(d1, d2);                                              // maxDist :169
    case 2: return (d1 == 0 || d2 == 0) ? max(d1, d2) : min(d1, d2);         // maxDistUnlessBlack :174
    case 3: return s1 < s2 ? d1 : d2;                                        // betterScore :182
    case 4: return s2 < 50 ? d2 : 0;                                         // onlyGood1 :190
    case 5: return (s1 < 100 && s2 < 100) ? (d1 + d2) / 2 : 0;               // onlyGoodAvg :198
    case 6: return (s1 < s2 && s1 < 20) ? 150 : ((s2 < s1 && s2 < 20) ? 255 : 0);  // overlap :205
    case 7: return 255 - s1;                                                 // blackToWhite :215
    default: return 0;
  }
}

constexpr int kTile = 32;

__global__ void __launch_bounds__(256) fuse_merge_kernel(const __grid_constant__ FuseArgs a) {
  __shared__ uint8_t t_d2[kTile][kTile + 4];
  __shared__ uint8_t t_s2[kTile][kTile + 4];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int i0 = blockIdx.y * kTile, j0 = blockIdx.x * kTile;

  // stage map/score 2: source row H-1-(x2+j), source column y2+i; lanes run along i (contiguous in memory)
  if (i0 < a.n && j0 < a.n) {
#pragma unroll
    for (int m = 0; m < kTile / 8; ++m) {
      const int j = j0 + ty + 8 * m, i = i0 + tx;
      if (j < a.n && i < a.n) {
        const size_t off = (size_t)(a.height - 1 - (a.x2 + j)) * a.step + (size_t)(a.y2 + i);
        t_d2[ty + 8 * m][tx] = a.d2[off];
        if (!a.scores_cropped) t_s2[ty + 8 * m][tx] = a.s2[off];
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int m = 0; m < kTile / 8; ++m) {
    const int i = i0 + ty + 8 * m, j = j0 + tx;
    if (i >= a.nc || j >= a.nc) continue;
    int out;
    if (i < a.n && j < a.n) {
      const size_t o1 = (size_t)(a.y1 + i) * a.step + (size_t)(a.x1 + j);
      const int d1 = a.d1[o1], d2 = t_d2[tx][ty + 8 * m];
      const int s1 = a.scores_cropped ? a.s1[(size_t)i * a.n + j] : a.s1[o1];
      const int s2 = a.scores_cropped ? a.s2[(size_t)i * a.n + j] : t_s2[tx][ty + 8 * m];
      out = fuse_rule(a.rule, d1, d2, s1, s2);
      // :118-121 -- score_1 == grad_1 == combined alias one buffer (:77, :113): min of the values read above
      if (a.combined) a.combined[(size_t)i * a.n + j] = (uint8_t)min(s1, s2);
    } else {
      // container pixels the merge loop never writes keep message 2's un-rotated pixels (:105-106)
      out = a.d2[(size_t)(a.yc + i) * a.step + (size_t)(a.xc + j)];
    }
    a.container[(size_t)i * a.nc + j] = (uint8_t)out;
  }
}

// ---- debug colouriser: DepthMapFusion::colorizeDepth (src/depth_map_fusion.cpp:304-358) -----------------------
// A pure function of the gray value, so a 256-entry table is built once on the device with the reference's float32
// operation order (explicit roundings: nothing may be contracted into an FMA) and applied per pixel.
__global__ void colorize_lut_kernel(uchar4 *lut) {
  const int g = threadIdx.x;
  const unsigned char d = (unsigned char)__double2int_rz(__dadd_rn(40.0, __dmul_rn(0.8, (double)g)));
  const unsigned int H = 255 - (255 - d) * 280 / 255;
  const unsigned int hi = (H / 60) % 6;
  const float f = __fsub_rn(__fdiv_rn((float)H, 60.f), (float)(H / 60));
  const float V = 1.f, p = 0.f;                       // S = V = 1: p = V*(1-S)
  const float q = __fsub_rn(1.f, f);                  // V*(1 - f*S)
  const float t = __fsub_rn(1.f, __fsub_rn(1.f, f));  // V*(1 - (1-f)*S)
  float rx = 0.f, ry = 0.f, rz = 0.f;
  if (hi == 0) rx = p, ry = t, rz = V;
  if (hi == 1) rx = p, ry = V, rz = q;
  if (hi == 2) rx = t, ry = V, rz = p;
  if (hi == 3) rx = V, ry = q, rz = p;
  if (hi == 4) rx = V, ry = p, rz = t;
  if (hi == 5) rx = q, ry = p, rz = V;
  auto to_u8 = [](float v) { return (unsigned char)__float2int_rz(__fmul_rn(fmaxf(0.f, fminf(v, 1.f)), 255.f)); };
  uchar4 o = make_uchar4(to_u8(rx), to_u8(ry), to_u8(rz), 0);
  if (d == 40) o = make_uchar4(0, 0, 0, 0);
  lut[g] = o;
}

__global__ void colorize_apply_kernel(const uint8_t *gray, size_t step, int w, int h, const uchar4 *lut, uint8_t *rgb) {
  __shared__ uchar4 s_lut[256];
  s_lut[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  const size_t n = (size_t)w * h;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
    const uchar4 c = s_lut[gray[(size_t)y * step + x]];
    rgb[3 * i + 0] = c.x, rgb[3 * i + 1] = c.y, rgb[3 * i + 2] = c.z;
  }
}

void crop_to_square(int cols, int rows, int ox, int oy, int member_oy, int r[4]) {
  const int num_cols = cols - std::abs(ox), num_rows = rows - std::abs(oy);
  const int n = std::min(cols, rows) - std::max(std::abs(ox), std::abs(member_oy));  // :252-253 uses offset_y_
  int sc, sr;
  if (num_cols < num_rows) {
    sc = std::max(0, ox);
    sr = std::max(0, oy + (num_rows - num_cols) / 2);
  } else {
    sc = std::max(0, ox + (num_cols - num_rows) / 2);
    sr = std::max(0, oy);
  }
  r[0] = sc, r[1] = sr, r[2] = n, r[3] = n;
}
bool rect_inside(const int r[4], int cols, int rows) {
  return r[2] >= 0 && r[0] >= 0 && r[1] >= 0 && r[0] + r[2] <= cols && r[1] + r[3] <= rows;
}

}  // namespace

bool fuse_geometry(int width, int height, int ox, int oy, int crop_l, int crop_r, int crop_t, int crop_b,
                   FuseGeometry *g) {
  crop_to_square(width, height, ox, oy, oy, g->r1);     // :48, :66
  crop_to_square(height, width, -ox, -oy, oy, g->r2);   // :57, :85 (rotated frame: cols = height)
  crop_to_square(width, height, 0, 0, oy, g->rc);       // :106
  g->n = g->r1[2];
  g->nc = g->rc[2];
  g->out_x = crop_l;
  g->out_y = crop_t;
  g->out_w = g->nc - crop_l - crop_r;
  g->out_h = g->nc - crop_t - crop_b;
  if (!rect_inside(g->r1, width, height) || !rect_inside(g->r2, height, width) || !rect_inside(g->rc, width, height))
    return false;
  if (g->n != g->r2[2] || g->n > g->nc || g->n <= 0) return false;
  if (g->out_w <= 0 || g->out_h <= 0 || crop_l < 0 || crop_t < 0 || crop_r < 0 || crop_b < 0) return false;
  return true;
}

cudaError_t launch_fuse_merge(const FuseLaunch &L, cudaStream_t stream, int *launches) {
  FuseArgs a{};
  a.d1 = L.d1, a.d2 = L.d2, a.s1 = L.s1, a.s2 = L.s2;
  a.step = L.step;
  a.width = L.width, a.height = L.height;
  a.x1 = L.g.r1[0], a.y1 = L.g.r1[1];
  a.x2 = L.g.r2[0], a.y2 = L.g.r2[1];
  a.xc = L.g.rc[0], a.yc = L.g.rc[1];
  a.n = L.g.n, a.nc = L.g.nc;
  a.rule = L.rule;
  a.scores_cropped = L.scores_cropped ? 1 : 0;
  a.container = L.container;
  a.combined = L.combined;
  const dim3 grid((a.nc + kTile - 1) / kTile, (a.nc + kTile - 1) / kTile);
  fuse_merge_kernel<<<grid, 256, 0, stream>>>(a);
  if (launches) *launches = 1;
  return cudaGetLastError();
}

}  // namespace d2pc

namespace d2pc {
cudaError_t launch_colorize(const uint8_t *gray, size_t step, int w, int h, void *lut256x4, bool build_lut,
                            uint8_t *rgb, cudaStream_t stream, int *launches) {
  uchar4 *lut = static_cast<uchar4 *>(lut256x4);
  int n = 0;
  if (build_lut) colorize_lut_kernel<<<1, 256, 0, stream>>>(lut), ++n;
  const size_t px = (size_t)w * h;
  const int grid = (int)std::min<size_t>((px + 255) / 256, 148 * 8);
  if (px) colorize_apply_kernel<<<grid, 256, 0, stream>>>(gray, step, w, h, lut, rgb), ++n;
  if (launches) *launches = n;
  return cudaGetLastError();
}
}  // namespace d2pc
