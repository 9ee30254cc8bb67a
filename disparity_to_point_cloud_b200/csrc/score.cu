// score.cu -- depth_map_fusion's matching-score preprocessing on sm_100a, one fused tile kernel per callback.
//
// Replaces, for MatchingScoreCb1 (src/depth_map_fusion.cpp:64-80) and MatchingScoreCb2 (:82-99):
//   cv::GaussianBlur(score, grad, Size(13,13), 3.0)          :70-71 / :89-90
//   cv::Sobel(grad, grad, -1, 0|2, 2|0, 7, 0.03)             :72    / :91
//   threshold(grad, grad, 30, 255, THRESH_BINARY)            :73    / :92
//   cv::GaussianBlur(grad, grad, Size(21,21), 10.0)          :74-75 / :93-94
//   grad = score + 2 * grad  (saturating)                    :76    / :95
// and the rotateMat + cropToSquare in front of callback 2 (:84-85) as index arithmetic.
//
// The arithmetic is OpenCV's, pinned against cv2 4.13.0 (tests/golden/score_chain_golden.npz, sepfilter_golden.npz):
//   * The FIRST blur runs on cropped_score_k_ = mat(region) (:264), a SUBMATRIX with the default non-isolated
//     border.  OpenCV 4.x keeps such a source out of its fixed-point Gaussian and runs sepFilter2D with the float32
//     kernel of getGaussianKernel(13, 3): float32 row pass (taps 0..12 in order), float32 symmetric column pass,
//     round half even.  Its AVX2 vector loops fuse multiply-add, the scalar tails do not (row pass: columns
//     >= n - n % 32, column pass: columns >= n - n % 4).  The border is the frame around the crop, reflect-101 only
//     at the frame's own edge.  (When the crop is the whole frame it is not a submatrix and takes the fixed-point
//     path like the second blur.)
//   * Sobel ksize 7 is the same float32 separable engine; the smoothing kernel carries the 0.03 scale.
//   * The SECOND blur runs on a stand-alone Mat: OpenCV's fixed-point Gaussian (kernel quantised to 8 fractional
//     bits summing to 256, rows 8.8, columns 16.16, round half up), reflect-101 at the n x n edge.
//
// One CTA produces a 64 x 64 output tile from the 102 x 102 frame pixels under it (halo 6 + 3 + 10), every
// intermediate in shared memory (55 KB): four global intermediates and six launches of the first version are gone.
// An intermediate array is indexed by EXTENDED crop coordinates and holds stage(reflect101(coordinate)), so the
// reflect-101 border of the stand-alone stages is a plain array read.
#include "score.h"

namespace d2pc {
namespace {

constexpr int kT = 64;              // output tile
constexpr int kHE = 10;             // halo of the thresholded edge image (second blur, 21 taps)
constexpr int kHG = kHE + 3;        // + Sobel 7
constexpr int kHF = kHG + 6;        // + first blur 13
constexpr int kEW = kT + 2 * kHE;   // 84
constexpr int kGW = kT + 2 * kHG;   // 90
constexpr int kFW = kT + 2 * kHF;   // 102
constexpr int kFP = 104, kGP = 92;  // byte pitches of the u8 arrays F and G (E reuses F's storage with pitch kEW)
constexpr int kThreads = 1024;
constexpr size_t kSmemF = (size_t)kFW * kFP;                // 10,608: frame tile, later the edge image E (84 x 84)
constexpr size_t kSmemR = (size_t)kFW * kGW * sizeof(float);  // 36,720: row-pass results (float / u16)
constexpr size_t kSmemG = (size_t)kGW * kGP;                // 8,280: first blur, u8
constexpr size_t kSmemS = (size_t)kT * kT;                  // 4,096: the score pixels of the tile
constexpr size_t kSmemTotal = kSmemF + kSmemR + kSmemG + kSmemS;

__constant__ float c_gauss13f[13] = {0x1.2fd344p-6f, 0x1.17e546p-5f, 0x1.cd7846p-5f, 0x1.54699ep-4f, 0x1.c16904p-4f,
                                     0x1.09752ep-3f, 0x1.189f6cp-3f, 0x1.09752ep-3f, 0x1.c16904p-4f, 0x1.54699ep-4f,
                                     0x1.cd7846p-5f, 0x1.17e546p-5f, 0x1.2fd344p-6f};  // getGaussianKernel(13, 3, CV_32F)
__constant__ int c_gauss13[13] = {5, 8, 15, 21, 28, 33, 36, 33, 28, 21, 15, 8, 5};  // fixed-point path
__constant__ int c_gauss21[21] = {9, 9, 11, 11, 12, 13, 13, 14, 14, 15, 14, 15, 14, 14, 13, 13, 12, 11, 11, 9, 9};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

// A reflected coordinate that some output of the tile really needs always lies within `halo` of the tile; array
// positions far past the image edge (a partial last tile) reflect further away, feed nothing, and are clamped so
// that their reads stay inside the arrays.
__device__ __forceinline__ int in_tile(int r, int t0, int halo) { return min(max(r, t0 - halo), t0 + kT - 1 + halo); }

struct ScoreArgs {
  const uint8_t *frame;  // stored frame (w x h, `step` bytes per row)
  size_t step;
  int w, h;
  int rotated;     // the chain runs on rotateMat(frame): rot(r, c) = frame(h - 1 - c, r); h cols x w rows
  int rx, ry, n;   // cropToSquare rectangle in the (rotated) frame
  int submatrix;   // the rectangle is smaller than the (rotated) frame: first blur = sepFilter2D (float32)
  float ks[7];     // Sobel smoothing kernel {1,6,15,20,15,6,1} * 0.03f, computed in float32 on the host
  uint8_t *out;    // n x n dense
};

__global__ void __launch_bounds__(kThreads) score_tile_kernel(const __grid_constant__ ScoreArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t *F = smem;                                        // [kFW][kFP]
  float *R = reinterpret_cast<float *>(smem + kSmemF);      // row-pass scratch
  uint8_t *G = smem + kSmemF + kSmemR;                      // [kGW][kGP]
  uint8_t *S = G + kSmemG;                                  // [kT][kT]
  uint8_t *E = F;                                           // [kEW][kEW], after F is dead
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kThreads / 32;
  const int n = a.n;
  const int tx0 = blockIdx.x * kT, ty0 = blockIdx.y * kT;   // crop coordinates of the tile
  const int cols = a.rotated ? a.h : a.w, rows = a.rotated ? a.w : a.h;  // of the (rotated) frame
  // every stage below: a warp takes a row of the stage's array, its lanes the columns (no index divisions)

  // ---- F: the frame under the tile, extended crop coordinates [t0 - 19, t0 + 83); non-isolated border =
  // the pixels around the crop, reflect-101 at the frame edge.  For the rotated view a warp takes a COLUMN of F:
  // its lanes walk down the rotated rows, i.e. along a stored row (coalesced), and the tile is written transposed.
  for (int u = warp; u < kFW; u += kWarps) {
    if (!a.rotated) {
      const uint8_t *row = a.frame + (size_t)reflect101(a.ry + ty0 - kHF + u, rows) * a.step;
      for (int v = lane; v < kFW; v += 32) F[u * kFP + v] = row[reflect101(a.rx + tx0 - kHF + v, cols)];
    } else {
      const uint8_t *row = a.frame + (size_t)(a.h - 1 - reflect101(a.rx + tx0 - kHF + u, cols)) * a.step;
      for (int v = lane; v < kFW; v += 32) F[v * kFP + u] = row[reflect101(a.ry + ty0 - kHF + v, rows)];
    }
  }
  __syncthreads();
  // the score pixels of the tile itself (for the last step; F is overwritten before that)
  for (int y = warp; y < kT; y += kWarps)
    for (int x = lane; x < kT; x += 32) S[y * kT + x] = F[(y + kHF) * kFP + x + kHF];

  // ---- first blur -> G(gy, gx) for extended coordinates [t0 - 13, t0 + 77)
  if (a.submatrix) {
    // rows: R(fy, gc) over all F rows and the G columns; taps in order, FMA in the vector columns
    const int row_vec_end = n - n % 32, col_vec_end = n - n % 4;
    // (a thread takes four adjacent output columns of one row: 16 pixels are loaded as four words and converted
    // once, instead of 13 loads + 13 conversions per output; every output still runs its own tap sequence)
    constexpr int kG4 = (kGW + 3) / 4;  // 23 groups of four columns
    for (int item = (int)threadIdx.x; item < kFW * kG4; item += kThreads) {
      const int fy = item / kG4, g = item - fy * kG4;
      const uint32_t *pw = reinterpret_cast<const uint32_t *>(F + fy * kFP + 4 * g);  // F column of crop column c0 - 6
      float v[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w = pw[j];
        v[4 * j + 0] = (float)(w & 0xffu), v[4 * j + 1] = (float)((w >> 8) & 0xffu);
        v[4 * j + 2] = (float)((w >> 16) & 0xffu), v[4 * j + 3] = (float)(w >> 24);
      }
      const int c0 = tx0 - kHG + 4 * g;  // crop column of the first output (the tail rule only matters for 0 <= c < n)
      float o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool fma = c0 + q < row_vec_end;
        float acc = __fmul_rn(c_gauss13f[0], v[q]);
#pragma unroll
        for (int k = 1; k < 13; ++k)
          acc = fma ? __fmaf_rn(c_gauss13f[k], v[q + k], acc) : __fadd_rn(acc, __fmul_rn(c_gauss13f[k], v[q + k]));
        o[q] = acc;
      }
      float *dst = R + fy * kGW + 4 * g;
      *reinterpret_cast<float2 *>(dst) = make_float2(o[0], o[1]);
      if (4 * g + 2 < kGW) *reinterpret_cast<float2 *>(dst + 2) = make_float2(o[2], o[3]);
    }
    __syncthreads();
    for (int gy = warp; gy < kGW; gy += kWarps) {
      const int ry = in_tile(reflect101(ty0 - kHG + gy, n), ty0, kHG);
      for (int gx = lane; gx < kGW; gx += 32) {
        const int rx = in_tile(reflect101(tx0 - kHG + gx, n), tx0, kHG);
        const float *t = R + (ry - (ty0 - kHF)) * kGW + (rx - (tx0 - kHG));  // R row of crop row ry, column rx
        const bool fma = rx < col_vec_end;
        float s = __fmul_rn(c_gauss13f[6], t[0]);
#pragma unroll
        for (int j = 1; j <= 6; ++j) {
          const float v = __fadd_rn(t[j * kGW], t[-j * kGW]);
          s = fma ? __fmaf_rn(c_gauss13f[6 + j], v, s) : __fadd_rn(s, __fmul_rn(c_gauss13f[6 + j], v));
        }
        G[gy * kGP + gx] = (uint8_t)fminf(fmaxf(rintf(s), 0.f), 255.f);
      }
    }
  } else {
    // the crop is the whole frame: fixed-point Gaussian (8.8 rows, 16.16 columns, round half up)
    uint32_t *R32 = reinterpret_cast<uint32_t *>(R);
    for (int fy = warp; fy < kFW; fy += kWarps)
      for (int gc = lane; gc < kGW; gc += 32) {
        const uint8_t *p = F + fy * kFP + gc;
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 13; ++k) acc += (uint32_t)c_gauss13[k] * p[k];
        R32[fy * kGW + gc] = acc;
      }
    __syncthreads();
    for (int gy = warp; gy < kGW; gy += kWarps) {
      const int ry = in_tile(reflect101(ty0 - kHG + gy, n), ty0, kHG);
      for (int gx = lane; gx < kGW; gx += 32) {
        const int rx = in_tile(reflect101(tx0 - kHG + gx, n), tx0, kHG);
        const uint32_t *t = R32 + (ry - (ty0 - kHF)) * kGW + (rx - (tx0 - kHG));
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 13; ++k) acc += (uint32_t)c_gauss13[k] * t[(k - 6) * kGW];
        G[gy * kGP + gx] = (uint8_t)min((acc + 32768u) >> 16, 255u);
      }
    }
  }
  __syncthreads();

  // ---- Sobel second derivative, ksize 7, x 0.03: row pass R(gy, ex), then column pass + saturate + threshold
  // -> E(ey, ex) for extended coordinates [t0 - 10, t0 + 74).  (rows of R are the G rows, which already hold
  // G(reflect(row)); columns are read at reflect(ex) + tap through the same convention.)
  {
    const int row_vec_end = n - n % 32, col_vec_end = n - n % 4;
    for (int gy = warp; gy < kGW; gy += kWarps)
      for (int ex = lane; ex < kEW; ex += 32) {
        const int rx = in_tile(reflect101(tx0 - kHE + ex, n), tx0, kHE);
        const uint8_t *p = G + gy * kGP + (rx - 3 - (tx0 - kHG));
        float s;
        if (!a.rotated) {  // dx = 0, dy = 2: scaled smoothing along x, taps in order, FMA in the vector columns
          const bool fma = rx < row_vec_end;
          s = __fmul_rn(a.ks[0], (float)p[0]);
#pragma unroll
          for (int k = 1; k < 7; ++k) {
            const float v = (float)p[k];
            s = fma ? __fmaf_rn(a.ks[k], v, s) : __fadd_rn(s, __fmul_rn(a.ks[k], v));
          }
        } else {  // dx = 2, dy = 0: integer second derivative along x (exact)
          const float d[7] = {1.f, 2.f, -1.f, -4.f, -1.f, 2.f, 1.f};
          s = 0.f;
#pragma unroll
          for (int k = 0; k < 7; ++k) s = __fadd_rn(s, __fmul_rn(d[k], (float)p[k]));
        }
        R[gy * kEW + ex] = s;
      }
    __syncthreads();
    for (int ey = warp; ey < kEW; ey += kWarps) {
      const int ry = in_tile(reflect101(ty0 - kHE + ey, n), ty0, kHE);
      for (int ex = lane; ex < kEW; ex += 32) {
        const int rx = reflect101(tx0 - kHE + ex, n);
        const float *t = R + (ry - (ty0 - kHG)) * kEW + ex;  // column ex already stands for reflect(ex)
        float s;
        if (!a.rotated) {
          const float d[4] = {-4.f, -1.f, 2.f, 1.f};  // centre, +-1, +-2, +-3 of {1,2,-1,-4,-1,2,1}
          s = __fmul_rn(d[0], t[0]);
#pragma unroll
          for (int j = 1; j <= 3; ++j) s = __fadd_rn(s, __fmul_rn(d[j], __fadd_rn(t[j * kEW], t[-j * kEW])));
        } else {
          const bool fma = rx < col_vec_end;
          s = __fmul_rn(a.ks[3], t[0]);
#pragma unroll
          for (int j = 1; j <= 3; ++j) {
            const float v = __fadd_rn(t[j * kEW], t[-j * kEW]);
            s = fma ? __fmaf_rn(a.ks[3 + j], v, s) : __fadd_rn(s, __fmul_rn(a.ks[3 + j], v));
          }
        }
        const float r = fminf(fmaxf(rintf(s), 0.f), 255.f);
        E[ey * kEW + ex] = r > 30.f ? 255 : 0;  // F is dead since the first blur's row pass
      }
    }
  }
  __syncthreads();

  // ---- second blur (fixed point) on E, then out = sat(score + 2 g)
  {
    // rows: four adjacent outputs per thread, the 21 taps as six 4-way byte dot products (DP4A) on the (shifted)
    // words of E; the 8.8 row sums go into Bt TRANSPOSED ([column][row], pitch 86) so that the column pass finds its
    // 21 vertical taps as consecutive 16-bit pairs and runs them as eleven 2-way dot products (DP2A).
    constexpr int kBP = 86;  // u16 per Bt column: 84 rows + 2 (43 words: odd, so adjacent columns fall in different banks)
    uint16_t *Bt = reinterpret_cast<uint16_t *>(R);  // [kT][kBP]
    uint32_t k4[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      k4[j] = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (4 * j + b < 21) k4[j] |= (uint32_t)c_gauss21[4 * j + b] << (8 * b);
    }
    for (int item = (int)threadIdx.x; item < (kT / 4) * kEW; item += kThreads) {
      const int g = item / kEW, ey = item - g * kEW;  // rows fastest: conflict-free loads (row pitch 21 words) and stores
      const uint32_t *pw = reinterpret_cast<const uint32_t *>(E + ey * kEW + 4 * g);  // E column of output 4g, tap 0
      uint32_t w[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) w[j] = pw[j];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t acc = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const uint32_t sw = q == 0 ? w[j] : (j < 5 ? __funnelshift_r(w[j], w[j + 1], 8 * q) : (w[5] >> (8 * q)));
          acc = __dp4a(sw, k4[j], acc);
        }
        Bt[(4 * g + q) * kBP + ey] = (uint16_t)acc;  // <= 255 * 256
      }
    }
    __syncthreads();
    // columns: output row oy needs Bt rows oy .. oy + 20; pairs are taken from the even row below oy, so the taps sit
    // at pair offset p = oy & 1 and the weights are shifted to match (a zero weight on the unused half)
    for (int oy = warp; oy < kT; oy += kWarps) {
      const int y = ty0 + oy;
      if (y >= n) break;
      const int par = oy & 1;
      for (int ox = lane; ox < kT; ox += 32) {
        const int x = tx0 + ox;
        if (x >= n) continue;
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(Bt + ox * kBP + (oy - par));
        uint32_t acc = 0;
#pragma unroll
        for (int j = 0; j < 11; ++j) {
          const int i0 = 2 * j - par, i1 = i0 + 1;  // taps of the two halves of word j
          const uint32_t wlo = (i0 >= 0 && i0 < 21) ? (uint32_t)c_gauss21[i0] : 0u;
          const uint32_t whi = (i1 >= 0 && i1 < 21) ? (uint32_t)c_gauss21[i1] : 0u;
          acc = __dp2a_lo(pw[j], wlo | (whi << 8), acc);
        }
        const uint32_t gq = min((acc + 32768u) >> 16, 255u);
        a.out[(size_t)y * n + x] = (uint8_t)min((uint32_t)S[oy * kT + ox] + 2u * gq, 255u);
      }
    }
  }
}

}  // namespace

cudaError_t launch_score_preprocess(const ScoreLaunch &L, cudaStream_t stream, int *launches) {
  const int n = L.rect[2];
  if (launches) *launches = 0;
  if (n <= 0 || L.rect[3] != n) return cudaErrorInvalidValue;
  ScoreArgs a{};
  a.frame = L.frame;
  a.step = L.step;
  a.w = L.width, a.h = L.height;
  a.rotated = L.rotated ? 1 : 0;
  a.rx = L.rect[0], a.ry = L.rect[1], a.n = n;
  const int cols = L.rotated ? L.height : L.width, rows = L.rotated ? L.width : L.height;
  a.submatrix = (n < cols || n < rows) ? 1 : 0;  // Mat::Mat(const Mat&, const Rect&) sets SUBMATRIX_FLAG iff smaller
  const float smooth[7] = {1, 6, 15, 20, 15, 6, 1};
  for (int i = 0; i < 7; ++i) {
    volatile float p = smooth[i] * 0.03f;  // float32 product, as Mat::operator*=(double) does for CV_32F
    a.ks[i] = p;
  }
  a.out = L.out;
  cudaError_t e = cudaFuncSetAttribute(score_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
  if (e != cudaSuccess) return e;  // per device, cheap: set on every launch
  const dim3 grid((n + kT - 1) / kT, (n + kT - 1) / kT);
  score_tile_kernel<<<grid, kThreads, kSmemTotal, stream>>>(a);
  if (launches) *launches = 1;
  return cudaGetLastError();
}

}  // namespace d2pc
