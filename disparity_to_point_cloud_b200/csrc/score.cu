// score.cu -- depth_map_fusion's matching-score preprocessing on sm_100a.
//
// Replaces, for MatchingScoreCb1 (src/depth_map_fusion.cpp:64-80) and MatchingScoreCb2 (:82-99):
//   cv::GaussianBlur(score, grad, Size(13,13), 3.0)          :70-71 / :89-90
//   cv::Sobel(grad, grad, -1, 0|2, 2|0, 7, 0.03)             :72    / :91
//   threshold(grad, grad, 30, 255, THRESH_BINARY)            :73    / :92
//   cv::GaussianBlur(grad, grad, Size(21,21), 10.0)          :74-75 / :93-94
//   grad = score + 2 * grad  (saturating)                    :76    / :95
// and the rotateMat + cropToSquare in front of callback 2 (:84-85) as index arithmetic.
//
// The arithmetic is OpenCV's, pinned against cv2 4.13.0 (tests/golden/score_chain_golden.npz):
//   * GaussianBlur on CV_8U is fixed point: kernel quantised to 8 fractional bits (error-diffused so it sums to
//     256), rows -> 8.8, columns -> 16.16, round half up.  Border reflect-101.  The first blur runs on a
//     non-isolated ROI, i.e. it sees the frame around the crop and reflects only at the frame edge.
//   * Sobel ksize 7 is a separable float32 filter; the smoothing kernel carries the 0.03 scale (in float32).  The
//     pass with the scaled kernel rounds at every step and OpenCV's AVX2 code uses FMA in its vector loop but
//     mul+add in the scalar tail (row pass: 32-column vectors; symmetric column pass: 4-column vectors).
// These images are small (n = 705 at 1280x720); the kernels are one thread per output pixel, cache-served.
#include "score.h"

namespace d2pc {
namespace {

__constant__ int c_gauss13[13] = {5, 8, 15, 21, 28, 33, 36, 33, 28, 21, 15, 8, 5};
__constant__ int c_gauss21[21] = {9, 9, 11, 11, 12, 13, 13, 14, 14, 15, 14, 15, 14, 14, 13, 13, 12, 11, 11, 9, 9};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

struct Src {  // a mono8 image, optionally viewed through rotateMat (rot(r, c) = src(h-1-c, r))
  const uint8_t *p;
  size_t step;
  int w, h;  // of the stored image
  bool rotated;
  __device__ __forceinline__ int cols() const { return rotated ? h : w; }
  __device__ __forceinline__ int rows() const { return rotated ? w : h; }
  __device__ __forceinline__ int at(int x, int y) const {
    return rotated ? p[(size_t)(h - 1 - x) * step + y] : p[(size_t)y * step + x];
  }
};

// Score 2 is read through rotateMat: walking along a row of the rotated image walks down a column of the stored
// frame, so a row filter over it would make 13 uncoalesced reads per output.  The (n + 12)^2 neighbourhood of the
// crop (reflect-101 at the frame edge, as the non-isolated ROI sees it) is therefore materialised once, through a
// 32 x 32 shared-memory tile: reads run along the stored rows, writes along the rows of the rotated view.
__global__ void __launch_bounds__(256) rotcrop_kernel(Src s, int x0, int y0, int m, uint8_t *out) {
  __shared__ uint8_t tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int xx = blockIdx.x * 32 + ty + 8 * j, yy = blockIdx.y * 32 + tx;
    if (xx < m && yy < m) tile[ty + 8 * j][tx] = (uint8_t)s.at(reflect101(x0 + xx, s.cols()), reflect101(y0 + yy, s.rows()));
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int xx = blockIdx.x * 32 + tx, yy = blockIdx.y * 32 + ty + 8 * j;
    if (xx < m && yy < m) out[(size_t)yy * m + xx] = tile[tx][ty + 8 * j];
  }
}

// rows pass of the fixed-point Gaussian: out16[(yy) * ow + x], yy in [0, oh + 2r), source row y0 + yy - r
template <int KS>
__global__ void gauss_rows_kernel(Src s, int x0, int y0, int ow, int oh, uint16_t *out16) {
  constexpr int R = KS / 2;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y;
  if (x >= ow) return;
  const int *k = KS == 13 ? c_gauss13 : c_gauss21;
  const int sy = reflect101(y0 + yy - R, s.rows());
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < KS; ++i) acc += (uint32_t)k[i] * (uint32_t)s.at(reflect101(x0 + x + i - R, s.cols()), sy);
  out16[(size_t)yy * ow + x] = (uint16_t)acc;  // 8.8, <= 255 * 256
}

// columns pass; kAddScore: out = sat(score + 2 * g) with score read from the (rotated) frame rectangle
template <int KS, bool kAddScore>
__global__ void gauss_cols_kernel(const uint16_t *rows16, int ow, int oh, Src score, int x0, int y0, uint8_t *out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= ow) return;
  const int *k = KS == 13 ? c_gauss13 : c_gauss21;
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < KS; ++i) acc += (uint32_t)k[i] * (uint32_t)rows16[(size_t)(y + i) * ow + x];
  uint32_t v = min((acc + 32768u) >> 16, 255u);
  if constexpr (kAddScore) v = min((uint32_t)score.at(x0 + x, y0 + y) + 2u * v, 255u);
  out[(size_t)y * ow + x] = (uint8_t)v;
}

struct SobelK {
  float ks[7];  // smoothing kernel {1,6,15,20,15,6,1} * 0.03f, computed in float32 on the host
};

// Sobel rows pass.  vertical == false (dx = 0, dy = 2): scaled smoothing kernel, sequential k = 0..6, FMA for
// columns < n - n % 32.  vertical == true (dx = 2, dy = 0): integer 2nd-derivative kernel (exact).
__global__ void sobel_rows_kernel(const uint8_t *src, int n, bool vertical, SobelK K, float *t) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= n) return;
  const uint8_t *row = src + (size_t)y * n;
  float s;
  if (!vertical) {
    const bool fma = x < n - n % 32;
    s = __fmul_rn(K.ks[0], (float)row[reflect101(x - 3, n)]);
#pragma unroll
    for (int i = 1; i < 7; ++i) {
      const float v = (float)row[reflect101(x + i - 3, n)];
      s = fma ? __fmaf_rn(K.ks[i], v, s) : __fadd_rn(s, __fmul_rn(K.ks[i], v));
    }
  } else {
    const float d[7] = {1.f, 2.f, -1.f, -4.f, -1.f, 2.f, 1.f};
    s = 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i) s = __fadd_rn(s, __fmul_rn(d[i], (float)row[reflect101(x + i - 3, n)]));
  }
  t[(size_t)y * n + x] = s;
}

// Sobel columns pass (symmetric form) + saturate_cast<uchar> (round half even) + threshold(30, 255, BINARY).
__global__ void sobel_cols_thresh_kernel(const float *t, int n, bool vertical, SobelK K, uint8_t *out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= n) return;
  float s;
  if (!vertical) {
    const float d[4] = {-4.f, -1.f, 2.f, 1.f};  // centre, +-1, +-2, +-3 of {1,2,-1,-4,-1,2,1}; exact in float32
    s = __fmul_rn(d[0], t[(size_t)y * n + x]);
#pragma unroll
    for (int i = 1; i <= 3; ++i)
      s = __fadd_rn(s, __fmul_rn(d[i], __fadd_rn(t[(size_t)reflect101(y + i, n) * n + x],
                                                 t[(size_t)reflect101(y - i, n) * n + x])));
  } else {
    const bool fma = x < n - n % 4;
    s = __fmul_rn(K.ks[3], t[(size_t)y * n + x]);
#pragma unroll
    for (int i = 1; i <= 3; ++i) {
      const float v = __fadd_rn(t[(size_t)reflect101(y + i, n) * n + x], t[(size_t)reflect101(y - i, n) * n + x]);
      s = fma ? __fmaf_rn(K.ks[3 + i], v, s) : __fadd_rn(s, __fmul_rn(K.ks[3 + i], v));
    }
  }
  const float r = fminf(fmaxf(rintf(s), 0.f), 255.f);
  out[(size_t)y * n + x] = r > 30.f ? 255 : 0;
}

inline size_t al(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

size_t score_scratch_bytes(int n) {
  const size_t nn = (size_t)n * n;
  return al((size_t)(n + 20) * n * 2) + al(nn * 4) + 2 * al(nn) + 256;
}

cudaError_t launch_score_preprocess(const ScoreLaunch &L, cudaStream_t stream, int *launches) {
  const int n = L.rect[2];
  if (launches) *launches = 0;
  if (n <= 0) return cudaErrorInvalidValue;
  Src s{L.frame, L.step, L.width, L.height, L.rotated};
  SobelK K;
  const float smooth[7] = {1, 6, 15, 20, 15, 6, 1};
  for (int i = 0; i < 7; ++i) {
    volatile float p = smooth[i] * 0.03f;  // float32 product, as Mat::operator*=(double) does for CV_32F
    K.ks[i] = p;
  }
  const dim3 blk(128);
  const dim3 g13((n + 127) / 128, n + 12), g21((n + 127) / 128, n + 20), gn((n + 127) / 128, n);
  int n_launch = 6;
  if (L.rotated && (size_t)(n + 12) * (n + 12) <= (size_t)n * n * 4) {
    // (the float scratch is free until the Sobel pass)
    uint8_t *roi = reinterpret_cast<uint8_t *>(L.f32);
    const int m = n + 12;
    rotcrop_kernel<<<dim3((m + 31) / 32, (m + 31) / 32), dim3(32, 8), 0, stream>>>(s, L.rect[0] - 6, L.rect[1] - 6, m, roi);
    Src r{roi, (size_t)m, m, m, false};
    gauss_rows_kernel<13><<<g13, blk, 0, stream>>>(r, 6, 6, n, n, L.rows16);
    n_launch = 7;
  } else {
    gauss_rows_kernel<13><<<g13, blk, 0, stream>>>(s, L.rect[0], L.rect[1], n, n, L.rows16);
  }
  gauss_cols_kernel<13, false><<<gn, blk, 0, stream>>>(L.rows16, n, n, s, 0, 0, L.tmp8a);
  sobel_rows_kernel<<<gn, blk, 0, stream>>>(L.tmp8a, n, L.rotated, K, L.f32);
  sobel_cols_thresh_kernel<<<gn, blk, 0, stream>>>(L.f32, n, L.rotated, K, L.tmp8b);
  Src e{L.tmp8b, (size_t)n, n, n, false};
  gauss_rows_kernel<21><<<g21, blk, 0, stream>>>(e, 0, 0, n, n, L.rows16);
  gauss_cols_kernel<21, true><<<gn, blk, 0, stream>>>(L.rows16, n, n, s, L.rect[0], L.rect[1], L.out);
  if (launches) *launches = n_launch;
  return cudaGetLastError();
}

}  // namespace d2pc
