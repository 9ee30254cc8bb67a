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

template <int K>
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
      if (col_ok) dst[(size_t)y * a.dst_step + x0 + lane] = (uint8_t)med;
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
// Column-histogram variant: the same exact median with two histogram updates per thread and row.
// ---------------------------------------------------------------------------
// Every image column of the warp's footprint (32 outputs + 2R halo) keeps the histogram of its K most recent
// pixels; moving down one row changes each column histogram by one pixel out, one pixel in.  The window histogram
// of an output is the sum of K adjacent column histograms and is never materialised: the running median needs
//   below  = #{window pixels < med}: updated from the 2K pixels that left / entered (register compares), and
//   W(med) = window count of bin med: K shared-memory reads of neighbouring columns (conflict-free layout).
// The serial read-modify-write chain per row shrinks from 2K to 2 (4 for the lanes that own a halo column), the
// rest is independent loads + ALU.  Layout: counter (bin, column c) at byte (bin>>2)*256 + c*4 + (bin&3): a warp
// touching 32 consecutive columns hits 32 distinct banks whatever the bins are.
constexpr int kColWarps = 4;
constexpr int kColThreads = kColWarps * 32;
constexpr int kColHistBytes = 64 * 256;  // 64 bin groups x 64 column slots x 4 counters = 16 KB per warp

__device__ __forceinline__ uint32_t col_off(uint32_t bin) { return ((bin >> 2) << 8) | (bin & 3u); }

template <int K>
__global__ void __launch_bounds__(kColThreads) median_col_kernel(const __grid_constant__ MedianArgs a) {
  constexpr int R = K / 2;
  constexpr int kRank = (K * K) / 2;
  extern __shared__ __align__(16) uint8_t s_dyn[];  // [kColWarps][kColHistBytes] histograms, then the rings
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  uint8_t *hist = s_dyn + (size_t)wic * kColHistBytes;
  uint8_t(*ring)[kRingPitch] = reinterpret_cast<uint8_t(*)[kRingPitch]>(s_dyn + (size_t)kColWarps * kColHistBytes +
                                                                       (size_t)wic * K * kRingPitch);
  uint8_t *own_a = hist + 4 * lane;         // histogram column of footprint column `lane`
  uint8_t *own_b = hist + 4 * (32 + lane);  // ... and of halo column 32 + lane (lanes < 2R)
  const bool has_b = lane < 2 * R;

  for (uint32_t unit = blockIdx.x * kColWarps + wic; unit < a.total_units; unit += gridDim.x * kColWarps) {
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
    const int gx_a = clampi(x0 - R + lane, 0, a.width - 1);
    const int gx_b = clampi(x0 - R + 32 + lane, 0, a.width - 1);

    // ---- zero the column histograms, load the first window's rows, build the column histograms
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < kColHistBytes / (32 * 16); ++i)
      reinterpret_cast<uint4 *>(hist)[i * 32 + lane] = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int s = 0; s < K; ++s) {
      const uint8_t *row = src + (size_t)clampi(y_first - R + s, 0, a.height - 1) * a.src_step;
      ring[s][lane] = row[gx_a];
      if (has_b) ring[s][32 + lane] = row[gx_b];
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < K; ++s) {
      const uint32_t oa = col_off(ring[s][lane]);
      own_a[oa] = own_a[oa] + 1;
      if (has_b) {
        const uint32_t ob = col_off(ring[s][32 + lane]);
        own_b[ob] = own_b[ob] + 1;
      }
    }
    __syncwarp();
    // window count of one bin / of one 4-bin word for this lane's output: columns lane .. lane+K-1
    auto wcount = [&](int bin) {
      const uint8_t *p = own_a + col_off((uint32_t)bin);
      int c = 0;
#pragma unroll
      for (int dx = 0; dx < K; ++dx) c += p[4 * dx];
      return c;
    };
    // ---- initial median: 4 bins at a time, then bin by bin
    int med = 0, below = 0;
    {
      int g = 0;
      for (; g < 64; ++g) {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(own_a + (g << 8));
        uint32_t lo = 0, hi = 0;  // byte sums of the K words, two 16-bit lanes each
#pragma unroll
        for (int dx = 0; dx < K; ++dx) {
          const uint32_t w = p[dx];
          lo += w & 0x00ff00ffu;
          hi += (w >> 8) & 0x00ff00ffu;
        }
        const int s4 = (int)((lo & 0xffffu) + (lo >> 16) + (hi & 0xffffu) + (hi >> 16));
        if (below + s4 > kRank) break;
        below += s4;
      }
      med = g * 4;
      for (;;) {
        const int c = wcount(med);
        if (below + c > kRank) break;
        below += c;
        ++med;
      }
    }

    int slot = 0;  // ring slot holding the oldest window row
    for (int y = y_first; y < y_end; ++y) {
      if (col_ok) dst[(size_t)y * a.dst_step + x0 + lane] = (uint8_t)med;
      if (y + 1 >= y_end) break;
      const uint8_t *row = src + (size_t)clampi(y + 1 + R, 0, a.height - 1) * a.src_step;
      const uint32_t na = row[gx_a];
      const uint32_t nb = has_b ? row[gx_b] : 0u;
      // pixels leaving the window: this lane's output loses ring[slot][lane .. lane+K-1]
      int d_below = 0;
#pragma unroll
      for (int dx = 0; dx < K; ++dx) d_below -= ((int)ring[slot][lane + dx] < med) ? 1 : 0;
      {
        const uint32_t oa = col_off(ring[slot][lane]), ia = col_off(na);  // column histograms: one out, one in
        own_a[oa] = own_a[oa] - 1;
        own_a[ia] = own_a[ia] + 1;
        if (has_b) {
          const uint32_t ob = col_off(ring[slot][32 + lane]), ib = col_off(nb);
          own_b[ob] = own_b[ob] - 1;
          own_b[ib] = own_b[ib] + 1;
        }
      }
      __syncwarp();
      ring[slot][lane] = (uint8_t)na;
      if (has_b) ring[slot][32 + lane] = (uint8_t)nb;
      __syncwarp();
#pragma unroll
      for (int dx = 0; dx < K; ++dx) d_below += ((int)ring[slot][lane + dx] < med) ? 1 : 0;
      below += d_below;
      slot = (slot + 1 == K) ? 0 : slot + 1;
      // re-centre: invariant below <= kRank < below + W(med)
      while (below > kRank) {
        --med;
        below -= wcount(med);
      }
      for (;;) {
        const int c = wcount(med);
        if (below + c > kRank) break;
        below += c;
        ++med;
      }
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
cudaError_t launch_col(const MedianArgs &a, int sm_count, cudaStream_t s) {
  const size_t smem = (size_t)kColWarps * kColHistBytes + (size_t)kColWarps * K * kRingPitch;
  cudaError_t e = cudaFuncSetAttribute(median_col_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = ((uint64_t)a.total_units + kColWarps - 1) / kColWarps;
  const uint64_t cap = (uint64_t)sm_count * 3;
  median_col_kernel<K><<<(int)(ctas < cap ? ctas : cap), kColThreads, smem, s>>>(a);
  return cudaGetLastError();
}

template <int K>
cudaError_t launch_k(const MedianArgs &a, int grid, cudaStream_t s) {
  median_hist_kernel<K><<<grid, kThreads, 0, s>>>(a);
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
  const uint64_t ctas = (total + kWarps - 1) / kWarps;
  const uint64_t cap = (uint64_t)L.sm_count * 6;
  const int grid = (int)(ctas < cap ? ctas : cap);
  if (launches) *launches = 1;
  if (L.variant == 1) {
    switch (L.ksize) {
      case 3: return launch_col<3>(a, L.sm_count, stream);
      case 5: return launch_col<5>(a, L.sm_count, stream);
      case 7: return launch_col<7>(a, L.sm_count, stream);
      case 9: return launch_col<9>(a, L.sm_count, stream);
      case 11: return launch_col<11>(a, L.sm_count, stream);
      case 13: return launch_col<13>(a, L.sm_count, stream);
      default: return launch_col<15>(a, L.sm_count, stream);
    }
  }
  if (L.ksize == 3 && L.variant == 0 && L.oh <= 65535 && L.n_frames <= 65535) {
    median3_net_kernel<<<dim3((unsigned)((L.ow + 127) / 128), (unsigned)L.oh, L.n_frames), 128, 0, stream>>>(a);
    return cudaGetLastError();
  }
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
