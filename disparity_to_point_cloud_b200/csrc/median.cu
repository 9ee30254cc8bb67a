// median.cu -- exact KxK median of CV_8UC1 images on sm_100a, replicate border.
//
// Replaces cv::medianBlur(img, out, 11)  src/disparity_to_point_cloud.cpp:55-57
//      and cv::medianBlur(img, img, 3)   src/depth_map_fusion.cpp:124
// (the result is an order statistic of 8-bit data, so any exact algorithm is
// bit-identical to OpenCV's; SURVEY.md A.4 pins the border as replicate).
//
// Algorithm: sliding-histogram (Huang) median, one output column per thread,
// sliding DOWN a strip of rows.  Each warp owns 32 adjacent columns and keeps
//   * a 256-bin x 32-lane histogram of 8-bit counters in shared memory, laid
//     out so that lane L only ever touches bank L (conflict-free for any data),
//   * a ring of the K most recent input rows (32+K-1 bytes each).
// Moving one row down removes K pixels and adds K pixels per thread, then the
// running median walks a few bins.  Cost is O(K) per output, independent of
// how disordered the image is (uniform noise is not a worst case).
#include "median.h"

#include <cstdint>

#include "reproject_math.cuh"

namespace d2pc {
namespace {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kRingPitch = 48;  // >= 32 + 15 - 1, multiple of 4

struct MedianArgs {
  const uint8_t *src;
  uint8_t *dst;
  size_t src_step, dst_step, src_frame_stride, dst_frame_stride;
  int width, height;
  int ox0, oy0, ow, oh;  // output region (inside the image); dst is addressed with image coordinates
  int strip_rows, n_colblk, n_strip;
  uint32_t units_per_frame, total_units;
  // fused DisparityCb (kFuse): the median goes straight through cpp:61-75 instead of to dst
  float4 *points;
  size_t points_frame_stride;  // in points
  float scale;
  int zero_numer;
  QParams Q;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// Offset of 8-bit counter `bin` inside a lane's private histogram column (add 4*lane for the byte address):
// monotonic in `bin`, so window pixels are kept in the ring already transformed and compared in this domain.
__device__ __forceinline__ uint32_t hist_off(uint32_t bin) { return ((bin >> 2) << 7) | (bin & 3u); }

// below += delta when off < med_off, as exactly two instructions (compare, predicated add); the compiler's own
// rendering of `below += (off < med_off) ? delta : 0` is three (compare, add into a temporary, predicated move).
template <int kDelta>
__device__ __forceinline__ void bump_if_below(int &below, uint32_t off, uint32_t med_off) {
  asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p add.s32 %0, %0, %3;\n\t}" : "+r"(below) : "r"(off), "r"(med_off), "n"(kDelta));
}

// kFuse != 0: instead of storing the median byte, finish the reference's callback for that pixel -- x 1/8 (cpp:61),
// reprojectImageTo3D with the rectified exact arithmetic (cpp:63-64), PointXYZ{x, y, z, 1.0f} at its crop position
// (cpp:67-75): a warp's 32 points are one 512-byte store.  The intermediate median image and the second launch of
// the mono8 callback disappear; the FP64 / conversion work issues in the gaps of this LSU-bound kernel.
// kFuse: 0 store the median, 1 fused callback, 2 fused callback with zero numerators kept straight-line (the
// arithmetic of the kMathRect0Z kernels: ~3 % slower, so only for a Q with an integral principal point column).
template <int K, int kFuse>
__global__ void __launch_bounds__(kThreads) median_hist_kernel(const __grid_constant__ MedianArgs a) {
  constexpr int R = K / 2;
  constexpr int kRank = (K * K) / 2;
  __shared__ __align__(16) uint8_t s_hist[kWarps][256 * 32];
  __shared__ __align__(4) uint16_t s_ring[kWarps][K][kRingPitch];  // hist_off() of the K most recent rows
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  uint8_t *hist = s_hist[wic];
  uint8_t *hb = hist + 4 * lane;  // this lane's column: counter `bin` lives at hb[hist_off(bin)]
  uint16_t(*ring)[kRingPitch] = s_ring[wic];

  for (uint32_t unit = blockIdx.x * kWarps + wic; unit < a.total_units; unit += gridDim.x * kWarps) {
    const uint32_t f = unit / a.units_per_frame;
    const uint32_t rem = unit - f * a.units_per_frame;
    const int strip = rem / a.n_colblk;
    const int cb = rem - strip * a.n_colblk;
    const int x0 = a.ox0 + cb * 32;
    const int y_first = a.oy0 + strip * a.strip_rows;
    const int y_end = min(y_first + a.strip_rows, a.oy0 + a.oh);
    const uint8_t *src = a.src + (size_t)f * a.src_frame_stride;
    uint8_t *dst = a.dst + (size_t)f * a.dst_frame_stride;
    const bool col_ok = (x0 + lane) < (a.ox0 + a.ow);
    // fused: this lane's column numerator X = (double)(float)(u + q03), for the whole strip
    double xd = 0.0;
    bool xslow = false, xzero = false;
    float4 *points = nullptr;
    if constexpr (kFuse != 0) {
      xd = rect_axis_const(x0 + lane, a.Q.q03);
      xslow = rect_axis_slow_t<kFuse == 2>(xd) || a.Q.zd_slow;
      xzero = kFuse == 2 && rect_axis_zero(xd);
      points = a.points + (size_t)f * a.points_frame_stride + (x0 + lane - a.ox0);
    }

    // columns this lane fetches for every ring row (replicate border = clamp)
    const int gx_a = clampi(x0 - R + lane, 0, a.width - 1);
    const int gx_b = clampi(x0 - R + 32 + lane, 0, a.width - 1);

    // ---- zero the histogram (warp-cooperative, 16 B per store)
    __syncwarp();
#pragma unroll
    for (int i = 0; i < (256 * 32) / (32 * 16); ++i)
      reinterpret_cast<uint4 *>(hist)[i * 32 + lane] = make_uint4(0, 0, 0, 0);

    // ---- fill the ring with the window rows of the first output row
#pragma unroll
    for (int s = 0; s < K; ++s) {
      const uint8_t *row = src + (size_t)clampi(y_first - R + s, 0, a.height - 1) * a.src_step;
      ring[s][lane] = (uint16_t)hist_off(row[gx_a]);
      if (lane < K - 1) ring[s][32 + lane] = (uint16_t)hist_off(row[gx_b]);
    }
    __syncwarp();
#pragma unroll 1
    for (int s = 0; s < K; ++s) {
#pragma unroll
      for (int dx = 0; dx < K; ++dx) {
        const uint32_t off = ring[s][lane + dx];
        hb[off] = hb[off] + 1;
      }
    }
    // ---- initial median: walk 4 bins (one word) at a time, then bin by bin
    int med = 0, below = 0;
    {
      int w = 0;
      for (; w < 64; ++w) {
        const uint32_t word = reinterpret_cast<const uint32_t *>(hist)[w * 32 + lane];
        const int s4 = (word & 0xff) + ((word >> 8) & 0xff) + ((word >> 16) & 0xff) + (word >> 24);
        if (below + s4 > kRank) break;
        below += s4;
      }
      med = w * 4;
      while (below + (int)hb[hist_off(med)] <= kRank) {
        below += hb[hist_off(med)];
        ++med;
      }
    }
    uint32_t med_off = hist_off(med);

    int slot = 0;  // ring slot holding the oldest window row
    for (int y = y_first; y < y_end; ++y) {
      if constexpr (kFuse != 0) {
        if (col_ok) {
          const float disp = __fadd_rn(__fmul_rn((float)med, a.scale), 0.0f);  // convertTo(CV_32FC1, 1/8), cpp:61
          const double yd = rect_axis_const(y, a.Q.q13);
          bool slow;
          float4 p = reproject_exact_rectified<true, true, kFuse == 2>(
              a.Q, xd, yd, xslow || rect_axis_slow_t<kFuse == 2>(yd), xzero || (kFuse == 2 && rect_axis_zero(yd)), disp, slow);
          if (__builtin_expect(slow, 0)) p = reproject_exact_slow(a.Q.q, x0 + lane, y, disp);
          __stcs(points + (size_t)(y - a.oy0) * a.ow, p);
        }
      } else {
        if (col_ok) dst[(size_t)y * a.dst_step + x0 + lane] = (uint8_t)med;
      }
      if (y + 1 >= y_end) break;
      // prefetch the row entering the window
      const uint8_t *row = src + (size_t)clampi(y + 1 + R, 0, a.height - 1) * a.src_step;
      const uint32_t na = hist_off(row[gx_a]);
      const uint32_t nb = (lane < K - 1) ? hist_off(row[gx_b]) : 0u;
      // remove the oldest row.  The K ring entries are read up front: the compiler can not move a ring load
      // across a histogram store (both are shared memory), so reading them inside the update loop would put two
      // dependent shared-memory latencies on every update instead of one.
      uint32_t offs[K];
#pragma unroll
      for (int dx = 0; dx < K; ++dx) offs[dx] = ring[slot][lane + dx];
#pragma unroll
      for (int dx = 0; dx < K; ++dx) {
        hb[offs[dx]] = hb[offs[dx]] - 1;
        bump_if_below<-1>(below, offs[dx], med_off);
      }
      __syncwarp();
      ring[slot][lane] = (uint16_t)na;
      if (lane < K - 1) ring[slot][32 + lane] = (uint16_t)nb;
      __syncwarp();
      // add the new row
#pragma unroll
      for (int dx = 0; dx < K; ++dx) offs[dx] = ring[slot][lane + dx];
#pragma unroll
      for (int dx = 0; dx < K; ++dx) {
        hb[offs[dx]] = hb[offs[dx]] + 1;
        bump_if_below<1>(below, offs[dx], med_off);
      }
      slot = (slot + 1 == K) ? 0 : slot + 1;
      // re-centre: invariant below <= kRank < below + hist[med]
      while (below > kRank) {
        --med;
        below -= hb[hist_off(med)];
      }
      for (;;) {
        const int hm = hb[hist_off(med)];
        if (below + hm > kRank) break;
        below += hm;
        ++med;
      }
      med_off = hist_off(med);
    }
  }
}

// ---------------------------------------------------------------------------
// 3 x 3: a 19-exchange selection network per output (depth_map_fusion.cpp:124)
// ---------------------------------------------------------------------------
// For K = 3 the sliding histogram is all start-up cost (a 256-bin histogram for 9 pixels); the median of nine is
// found with 19 min / max exchanges instead.  One thread per output pixel, replicate border by clamping.
__device__ __forceinline__ void exch(int &a, int &b) {
  const int lo = min(a, b);
  b = max(a, b);
  a = lo;
}
__global__ void __launch_bounds__(128) median3_net_kernel(const __grid_constant__ MedianArgs a) {
  const int x = a.ox0 + (int)(blockIdx.x * 128 + threadIdx.x);
  const int y = a.oy0 + (int)blockIdx.y;
  if (x >= a.ox0 + a.ow) return;
  const uint8_t *src = a.src + (size_t)blockIdx.z * a.src_frame_stride;
  const int xm = max(x - 1, 0), xp = min(x + 1, a.width - 1);
  const uint8_t *r0 = src + (size_t)max(y - 1, 0) * a.src_step;
  const uint8_t *r1 = src + (size_t)y * a.src_step;
  const uint8_t *r2 = src + (size_t)min(y + 1, a.height - 1) * a.src_step;
  int p0 = r0[xm], p1 = r0[x], p2 = r0[xp], p3 = r1[xm], p4 = r1[x], p5 = r1[xp], p6 = r2[xm], p7 = r2[x], p8 = r2[xp];
  exch(p1, p2), exch(p4, p5), exch(p7, p8);
  exch(p0, p1), exch(p3, p4), exch(p6, p7);
  exch(p1, p2), exch(p4, p5), exch(p7, p8);
  exch(p0, p3), exch(p5, p8), exch(p4, p7);
  exch(p3, p6), exch(p1, p4), exch(p2, p5);
  exch(p4, p7), exch(p4, p2), exch(p6, p4);
  exch(p4, p2);
  (a.dst + (size_t)blockIdx.z * a.dst_frame_stride)[(size_t)y * a.dst_step + x] = (uint8_t)p4;
}

template <int K>
cudaError_t launch_k(const MedianArgs &a, int grid, cudaStream_t s) {
  if (a.points && a.zero_numer) median_hist_kernel<K, 2><<<grid, kThreads, 0, s>>>(a);
  else if (a.points) median_hist_kernel<K, 1><<<grid, kThreads, 0, s>>>(a);
  else median_hist_kernel<K, 0><<<grid, kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_median_u8(const MedianLaunch &L, cudaStream_t stream, int *launches) {
  if (launches) *launches = 0;
  if (L.ow <= 0 || L.oh <= 0 || L.n_frames == 0 || L.width <= 0 || L.height <= 0) return cudaSuccess;
  if (L.ksize < 3 || L.ksize > 15 || (L.ksize & 1) == 0) return cudaErrorInvalidValue;
  MedianArgs a{};
  a.src = L.src;
  a.dst = L.dst;
  a.src_step = L.src_step;
  a.dst_step = L.dst_step;
  a.src_frame_stride = L.src_frame_stride;
  a.dst_frame_stride = L.dst_frame_stride;
  a.width = L.width;
  a.height = L.height;
  a.ox0 = L.ox0;
  a.oy0 = L.oy0;
  a.ow = L.ow;
  a.oh = L.oh;
  if (L.points) {
    if (!L.Q || L.points_stride_bytes % 16 != 0) return cudaErrorInvalidValue;
    a.points = reinterpret_cast<float4 *>(L.points);
    a.points_frame_stride = L.points_stride_bytes / 16;
    a.scale = L.scale;
    a.zero_numer = L.zero_numer ? 1 : 0;
    a.Q = *L.Q;
  }
  if (launches) *launches = 1;
  if (L.ksize == 3 && L.variant == 0 && !L.points && L.oh <= 65535 && L.n_frames <= 65535) {
    median3_net_kernel<<<dim3((unsigned)((L.ow + 127) / 128), (unsigned)L.oh, L.n_frames), 128, 0, stream>>>(a);
    return cudaGetLastError();
  }
  // one window histogram per output column (variant 2 forces it for ksize 3 as well).  Three other formulations were
  // built and measured slower in round 2 (experiments/median_variants.cu).
  a.n_colblk = (L.ow + 31) / 32;
  // strip height: long strips amortise the K*K start-up (throughput: batches), short ones fill the chip with
  // warps (latency: a single frame is fastest with 2-row strips, measured 35 vs 55 us at 752x480 -- every update
  // is a link of one dependent shared-memory chain, so a lone frame wants as many short chains as possible)
  int strip = 64;
  const uint64_t want = (uint64_t)L.sm_count * 16;
  while (strip > 2 && (uint64_t)a.n_colblk * ((L.oh + strip - 1) / strip) * L.n_frames < want) strip >>= 1;
  if (L.strip_rows > 0) strip = L.strip_rows;
  a.strip_rows = strip;
  a.n_strip = (L.oh + strip - 1) / strip;
  a.units_per_frame = (uint32_t)a.n_colblk * (uint32_t)a.n_strip;
  const uint64_t total = (uint64_t)a.units_per_frame * L.n_frames;
  if (total > 0xffffffffull) return cudaErrorInvalidValue;
  a.total_units = (uint32_t)total;
  const uint64_t cap = (uint64_t)L.sm_count * 6;
  const uint64_t ctas = (total + kWarps - 1) / kWarps;
  const int grid = (int)(ctas < cap ? ctas : cap);
  switch (L.ksize) {
    case 3: return launch_k<3>(a, grid, stream);
    case 5: return launch_k<5>(a, grid, stream);
    case 7: return launch_k<7>(a, grid, stream);
    case 9: return launch_k<9>(a, grid, stream);
    case 11: return launch_k<11>(a, grid, stream);
    case 13: return launch_k<13>(a, grid, stream);
    default: return launch_k<15>(a, grid, stream);
  }
}

}  // namespace d2pc
