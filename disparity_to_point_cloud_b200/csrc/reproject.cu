// reproject.cu -- the disparity callback's reprojection + crop + pack on sm_100a.
//
// Replaces, in ONE pass over HBM (4 B read + 16 B written per kept pixel):
//   cv::reprojectImageTo3D(real_disparity, image3D, Q_)   src/disparity_to_point_cloud.cpp:63-64
//   the 40-pixel border crop loop + pcl::PointXYZ push_back   :69-76
//   pcl::toROSMsg's memcpy into PointCloud2.data              :84-85
// and, for the mono8 entry, Mat::convertTo(CV_32FC1, 1/8)     :60-61.
//
// CROP kernel (reference-exact filter: output offsets are closed form)
//   work unit = 128 crop columns x RB crop rows, one warp per unit, grid-stride.
//   Per row: each lane issues one 16-byte streaming load (4 disparities), the
//   warp transposes through 512 B of shared memory so that lane L then owns
//   pixels L, L+32, L+64, L+96 -- which makes every one of the four 16-byte
//   point stores of the warp a contiguous 512-byte burst.  Column constants
//   (X) live in registers for the whole unit, row constants (Y) are computed
//   by one lane each and broadcast by shuffle; Q sits in the kernel parameter
//   (constant) bank.
//
// CROP_FINITE kernel (extension: drop non-finite points, keep row-major order)
//   single-pass chained scan with decoupled look-back: warp ballot/popc ->
//   block scan -> 64-bit {epoch,flag,value} tile descriptors -> compacted tile
//   staged in shared memory -> coalesced 16-byte stores.
#include "reproject.h"

#include "reproject_math.cuh"

namespace d2pc {

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kSegCols = 128;  // crop columns per warp-row: 32 lanes x 4

struct ReprojArgs {
  const uint8_t *in;
  size_t step, frame_stride;
  float4 *out;
  size_t out_frame_stride;  // in points
  uint32_t *counts;
  int n_frames, width, height, border, cw, ch;
  int rows_per_unit, n_seg, n_rb;
  uint32_t units_per_frame, total_units;
  float scale;
  // compaction
  unsigned long long *tile_desc;
  uint32_t *ticket;
  uint32_t epoch, tiles_per_frame;
  QParams Q;
};

enum { kMathRect0 = 0, kMathRectW = 1, kMathGeneric = 2, kMathFast = 3 };
#define D2PC_IS_RECT(m) ((m) == kMathRect0 || (m) == kMathRectW)

__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// src/disparity_to_point_cloud.cpp:61 -- convertTo(CV_32FC1, 1/8): float(src)*alpha + 0
__device__ __forceinline__ float u8_to_disp(uint32_t b, float scale) {
  return __fadd_rn(__fmul_rn((float)b, scale), 0.0f);
}

template <typename InT>
__device__ __forceinline__ float4 load4(const uint8_t *row, int col, float scale);
template <>
__device__ __forceinline__ float4 load4<float>(const uint8_t *row, int col, float) {
  return ld_stream_f4(reinterpret_cast<const float4 *>(row + (size_t)col * 4));
}
template <>
__device__ __forceinline__ float4 load4<uint8_t>(const uint8_t *row, int col, float scale) {
  const uint32_t w = __ldcs(reinterpret_cast<const uint32_t *>(row + col));
  return make_float4(u8_to_disp(w & 0xff, scale), u8_to_disp((w >> 8) & 0xff, scale),
                     u8_to_disp((w >> 16) & 0xff, scale), u8_to_disp(w >> 24, scale));
}
template <typename InT>
__device__ __forceinline__ float load1(const uint8_t *row, int col, float scale);
template <>
__device__ __forceinline__ float load1<float>(const uint8_t *row, int col, float) {
  return __ldcs(reinterpret_cast<const float *>(row) + col);
}
template <>
__device__ __forceinline__ float load1<uint8_t>(const uint8_t *row, int col, float scale) {
  return u8_to_disp(__ldcs(row + col), scale);
}

// Four pixels of one lane (columns u0, u0+32, u0+64, u0+96 of row v): the arithmetic is straight-line so the
// four FP64 dependency chains interleave; the rare slow path is one warp-level branch afterwards.
template <int kMath>
__device__ __forceinline__ void points_of4(const QParams &Q, const double (&xd)[4], double yd, uint32_t xslow,
                                           bool yslow, int u0, int v, const float (&d)[4], float4 (&p)[4]) {
  if constexpr (D2PC_IS_RECT(kMath)) {
    bool slow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      p[k] = reproject_exact_rectified<kMath == kMathRect0>(Q, xd[k], yd, yslow || ((xslow >> k) & 1u), d[k],
                                                            slow[k]);
    if (__builtin_expect(slow[0] || slow[1] || slow[2] || slow[3], 0)) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (slow[k]) p[k] = reproject_exact_slow(Q.q, u0 + 32 * k, v, d[k]);
    }
  } else if constexpr (kMath == kMathGeneric) {
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = reproject_exact_generic(Q.q, u0 + 32 * k, v, d[k]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = reproject_fast(Q.qf, u0 + 32 * k, v, d[k]);
  }
}

template <int kMath>
__device__ __forceinline__ float4 point_of(const QParams &Q, int u, int v, float d) {
  if constexpr (D2PC_IS_RECT(kMath)) {
    const double xd = rect_axis_const(u, Q.q03), yd = rect_axis_const(v, Q.q13);
    bool slow;
    float4 p = reproject_exact_rectified<kMath == kMathRect0>(
        Q, xd, yd, rect_axis_slow(xd) || rect_axis_slow(yd) || Q.zd_slow, d, slow);
    if (__builtin_expect(slow, 0)) p = reproject_exact_slow(Q.q, u, v, d);
    return p;
  } else if constexpr (kMath == kMathGeneric) {
    return reproject_exact_generic(Q.q, u, v, d);
  } else {
    return reproject_fast(Q.qf, u, v, d);
  }
}

// ---------------------------------------------------------------------------
// CROP kernel
// ---------------------------------------------------------------------------
template <typename InT, bool kVec, int kMath, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) reproject_crop_kernel(const __grid_constant__ ReprojArgs a) {
  __shared__ __align__(16) float stage[kWarpsPerCta][2][kSegCols];
  const int lane = threadIdx.x & 31;
  const int wic = threadIdx.x >> 5;
  const QParams &Q = a.Q;

  // one work unit (128 crop columns x rows_per_unit crop rows) per warp; the hardware CTA scheduler balances
  const uint32_t unit = blockIdx.x * kWarpsPerCta + wic;
  if (unit >= a.total_units) return;
  const uint32_t f = unit / a.units_per_frame;
  const uint32_t rem = unit - f * a.units_per_frame;
  const int rb = rem / a.n_seg;
  const int seg = rem - rb * a.n_seg;
  const int c_base = seg * kSegCols;
  const int r_base = rb * a.rows_per_unit;
  const int rows = min(a.rows_per_unit, a.ch - r_base);

  // column constants live in registers for the whole unit; row constants are computed by one lane per row
  double xd[4];
  uint32_t xslow = 0;
  double yd_lane = 0.0;
  if constexpr (D2PC_IS_RECT(kMath)) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      xd[k] = rect_axis_const(a.border + c_base + 32 * k + lane, Q.q03);
      xslow |= rect_axis_slow(xd[k]) ? (1u << k) : 0u;
    }
    if (Q.zd_slow) xslow = 0xfu;
    yd_lane = rect_axis_const(a.border + r_base + lane, Q.q13);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) xd[k] = 0.0;
  }

  const uint8_t *in_row = a.in + (size_t)f * a.frame_stride + (size_t)(a.border + r_base) * a.step;
  float4 *out_row = a.out + (size_t)f * a.out_frame_stride + (size_t)r_base * a.cw;

  for (int r = 0; r < rows; r += 2, in_row += 2 * a.step, out_row += 2 * (size_t)a.cw) {
    // ---- load two crop rows of this segment
    float dd[2][4];
    if constexpr (kVec) {
      const int c4 = c_base + 4 * lane;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (r + j < rows && c4 < a.cw) {
          const float4 v = load4<InT>(in_row + (size_t)j * a.step, a.border + c4, a.scale);
          *reinterpret_cast<float4 *>(&stage[wic][j][4 * lane]) = v;
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) dd[j][k] = stage[wic][j][32 * k + lane];
      __syncwarp();
    } else {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = c_base + 32 * k + lane;
          dd[j][k] = (r + j < rows && c < a.cw) ? load1<InT>(in_row + (size_t)j * a.step, a.border + c, a.scale) : 1.0f;
        }
      }
    }
    // ---- compute + store: lane L owns columns L, L+32, L+64, L+96 -> each store is a 512-byte burst
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (r + j >= rows) break;
      double yd = 0.0;
      bool yslow = false;
      if constexpr (D2PC_IS_RECT(kMath)) {
        yd = __shfl_sync(0xffffffffu, yd_lane, r + j);
        yslow = rect_axis_slow(yd);
      }
      float4 p[4];
      points_of4<kMath>(Q, xd, yd, xslow, yslow, a.border + c_base + lane, a.border + r_base + r + j, dd[j], p);
      float4 *o = out_row + (size_t)j * a.cw + c_base + lane;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (c_base + 32 * k + lane < a.cw) st_stream_f4(o + 32 * k, p[k]);
    }
  }
}

// ---------------------------------------------------------------------------
// CROP_FINITE kernel: order-preserving stream compaction, decoupled look-back
// ---------------------------------------------------------------------------
// Tile = kTilePts consecutive crop pixels in row-major order of one frame.
// Descriptor word: [63:34] launch epoch, [33:32] flag, [31:0] value.
constexpr int kCWarps = 8;
constexpr int kCThreads = kCWarps * 32;
constexpr int kItems = 8;                      // pixels per thread
constexpr int kTilePts = kCThreads * kItems;   // 2048
constexpr uint32_t kFlagAggregate = 1, kFlagPrefix = 2;

__device__ __forceinline__ unsigned long long desc_pack(uint32_t epoch, uint32_t flag, uint32_t value) {
  return ((unsigned long long)epoch << 34) | ((unsigned long long)flag << 32) | value;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <typename InT, bool kVec, int kMath>
__global__ void __launch_bounds__(kCThreads) reproject_compact_kernel(const __grid_constant__ ReprojArgs a) {
  __shared__ __align__(16) float4 tile_pts[kTilePts];  // 32 KB: the compacted tile
  __shared__ uint32_t warp_sums[kCWarps];
  __shared__ uint32_t s_tile, s_excl;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const QParams &Q = a.Q;
  const uint32_t pts_per_frame = (uint32_t)a.cw * (uint32_t)a.ch;
  const uint32_t total_tiles = a.tiles_per_frame * (uint32_t)a.n_frames;

  for (;;) {
    // dynamic tile id: a tile's predecessors have all started, so the look-back cannot deadlock
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= total_tiles) break;
    const uint32_t f = tile / a.tiles_per_frame;
    const uint32_t t_in_f = tile - f * a.tiles_per_frame;
    const uint8_t *in_f = a.in + (size_t)f * a.frame_stride;

    // ---- load + reproject kItems pixels per thread (blocked: thread t owns pixels first .. first+kItems-1)
    const uint32_t first = t_in_f * kTilePts + threadIdx.x * kItems;
    float dd[kItems];
    if constexpr (kVec) {  // cw % 4 == 0: a group of 4 never straddles a crop row
#pragma unroll
      for (int g = 0; g < kItems / 4; ++g) {
        const uint32_t idx = first + 4 * g;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < pts_per_frame) {
          const int crow = idx / (uint32_t)a.cw;
          const int c = idx - crow * a.cw;
          v = load4<InT>(in_f + (size_t)(a.border + crow) * a.step, a.border + c, a.scale);
        }
        dd[4 * g + 0] = v.x, dd[4 * g + 1] = v.y, dd[4 * g + 2] = v.z, dd[4 * g + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        const uint32_t idx = first + i;
        dd[i] = 0.f;
        if (idx < pts_per_frame) {
          const int crow = idx / (uint32_t)a.cw;
          const int c = idx - crow * a.cw;
          dd[i] = load1<InT>(in_f + (size_t)(a.border + crow) * a.step, a.border + c, a.scale);
        }
      }
    }
    float4 pts[kItems];
    uint32_t keep = 0;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const uint32_t idx = first + i;
      if (idx < pts_per_frame) {
        const int crow = idx / (uint32_t)a.cw;
        const int c = idx - crow * a.cw;
        const int u = a.border + c, v = a.border + crow;
        pts[i] = point_of<kMath>(Q, u, v, dd[i]);
        keep |= point_is_finite(pts[i]) ? (1u << i) : 0u;
      }
    }
    // ---- block-wide exclusive scan of per-thread counts
    const uint32_t cnt = __popc(keep);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) warp_sums[wic] = incl;
    __syncthreads();
    uint32_t warp_off = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kCWarps; ++w) {
      const uint32_t s = warp_sums[w];
      if (w < wic) warp_off += s;
      tile_total += s;
    }
    uint32_t local = warp_off + incl - cnt;

    // ---- decoupled look-back (warp 0); tile 0 of a frame starts the chain
    if (wic == 0) {
      unsigned long long *desc = a.tile_desc + tile;
      uint32_t excl = 0;
      if (t_in_f == 0) {
        if (lane == 0) st_relaxed_u64(desc, desc_pack(a.epoch, kFlagPrefix, tile_total));
      } else {
        if (lane == 0) st_relaxed_u64(desc, desc_pack(a.epoch, kFlagAggregate, tile_total));
        int look = (int)t_in_f - 1;  // predecessor window [look-31, look] within this frame
        for (;;) {
          const int my = look - lane;
          unsigned long long dv = 0;
          uint32_t flag = kFlagPrefix, val = 0;  // lanes before the frame start act as a zero prefix
          if (my >= 0) {
            do {
              dv = ld_relaxed_u64(a.tile_desc + (tile - t_in_f) + my);
            } while ((uint32_t)(dv >> 34) != a.epoch || ((dv >> 32) & 3u) == 0);
            flag = (uint32_t)(dv >> 32) & 3u;
            val = (uint32_t)dv;
          }
          const uint32_t pmask = __ballot_sync(0xffffffffu, flag == kFlagPrefix);
          // lanes nearer than (and including) the first PREFIX contribute
          const int stop = pmask ? (__ffs(pmask) - 1) : 32;
          uint32_t contrib = (lane <= stop) ? val : 0u;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
          excl += contrib;
          if (pmask) break;
          look -= 32;
        }
        if (lane == 0) st_relaxed_u64(desc, desc_pack(a.epoch, kFlagPrefix, excl + tile_total));
      }
      if (lane == 0) {
        s_excl = excl;
        if (t_in_f == a.tiles_per_frame - 1 && a.counts) a.counts[f] = excl + tile_total;
      }
    }
    // ---- compact into shared memory, then coalesced copy-out
#pragma unroll
    for (int i = 0; i < kItems; ++i)
      if (keep & (1u << i)) tile_pts[local++] = pts[i];
    __syncthreads();
    float4 *out_f = a.out + (size_t)f * a.out_frame_stride + s_excl;
    for (uint32_t i = threadIdx.x; i < tile_total; i += kCThreads) st_stream_f4(out_f + i, tile_pts[i]);
  }
}

template <typename InT, int kMath>
cudaError_t launch_typed(const ReprojArgs &a, bool vec, bool compact, int grid, int min_blocks, cudaStream_t s) {
  if (compact) {
    if (vec && a.cw % 4 == 0)
      reproject_compact_kernel<InT, true, kMath><<<grid, kCThreads, 0, s>>>(a);
    else
      reproject_compact_kernel<InT, false, kMath><<<grid, kCThreads, 0, s>>>(a);
  } else if (vec) {
    switch (min_blocks) {  // 128-thread CTAs: 4 -> <=128 regs, 6 -> 80, 7 -> 72, 8 -> 64
      case 4: reproject_crop_kernel<InT, true, kMath, 4><<<grid, kThreads, 0, s>>>(a); break;
      case 6: reproject_crop_kernel<InT, true, kMath, 6><<<grid, kThreads, 0, s>>>(a); break;
      case 8: reproject_crop_kernel<InT, true, kMath, 8><<<grid, kThreads, 0, s>>>(a); break;
      default: reproject_crop_kernel<InT, true, kMath, 7><<<grid, kThreads, 0, s>>>(a); break;
    }
  } else {
    reproject_crop_kernel<InT, false, kMath, 6><<<grid, kThreads, 0, s>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace

// Host: classify Q (disparity_to_point_cloud.hpp:90-104 produces the rectified form).
void make_qparams(const double q[16], QParams *out) {
  QParams P{};
  for (int i = 0; i < 16; ++i) {
    P.q[i] = q[i];
    P.qf[i] = (float)q[i];
  }
  auto bits = [](double x) {
    uint64_t b;
    memcpy(&b, &x, 8);
    return b;
  };
  auto pzero = [&](double x) { return bits(x) == 0; };
  auto fin = [](double x) { return x == x && x - x == 0.0; };
  const double a32 = q[14] < 0 ? -q[14] : q[14];
  const bool rect = q[0] == 1.0 && pzero(q[1]) && pzero(q[2]) && pzero(q[4]) && q[5] == 1.0 && pzero(q[6]) &&
                    pzero(q[8]) && pzero(q[9]) && pzero(q[10]) && pzero(q[12]) && pzero(q[13]) && fin(q[3]) &&
                    fin(q[7]) && fin(q[11]) && fin(q[15]) && fin(q[14]) && a32 > 1e-30 && a32 < 1e30;
  P.rectified = rect ? 1 : 0;
  P.q03 = q[3];
  P.q13 = q[7];
  P.q32 = q[14];
  P.q33 = q[15];
  P.q33_zero = (q[15] == 0.0) ? 1 : 0;
  // h2 = ((0*u + 0*v) + 0*d) + q23 = (+0) + q23 for finite u, v >= 0 and finite d
  volatile double z = 0.0;
  const double h2 = z + q[11];
  P.zd = (double)(float)h2;
  const uint64_t zb = bits(P.zd);
  P.zd_slow = ((zb << 1) == 0 || ((zb >> 52) & 0x7ff) == 0x7ff) ? 1 : 0;
  {
    uint32_t zi = (zb << 1) == 0 ? 0xFFC00000u : (uint32_t)(((zb >> 63) << 31) | 0x7f800000u);
    memcpy(&P.zinf, &zi, 4);
  }
  *out = P;
}

size_t reproject_scratch_bytes(uint32_t n_frames, uint32_t width, uint32_t height, int border) {
  const long cw = (long)width - 2L * border, ch = (long)height - 2L * border;
  if (cw <= 0 || ch <= 0) return 256;
  const uint64_t pts = (uint64_t)cw * ch;
  const uint64_t tiles = (pts + kTilePts - 1) / kTilePts;
  return (size_t)(tiles * n_frames * 8 + 256);
}

cudaError_t launch_reproject(const ReprojectLaunch &L, cudaStream_t stream, int *launches) {
  const long cw = (long)L.width - 2L * L.border, ch = (long)L.height - 2L * L.border;
  if (launches) *launches = 0;
  if (cw <= 0 || ch <= 0 || L.n_frames == 0) {
    // nothing kept: the reference publishes an empty cloud
    if (L.counts && L.n_frames) {
      if (launches) *launches = 0;
      return cudaMemsetAsync(L.counts, 0, sizeof(uint32_t) * L.n_frames, stream);
    }
    return cudaSuccess;
  }
  ReprojArgs a{};
  a.in = static_cast<const uint8_t *>(L.in);
  a.step = L.step;
  a.frame_stride = L.frame_stride;
  a.out = reinterpret_cast<float4 *>(L.out);
  a.out_frame_stride = L.out_stride_bytes / 16;
  a.counts = L.counts;
  a.n_frames = (int)L.n_frames;
  a.width = (int)L.width;
  a.height = (int)L.height;
  a.border = L.border;
  a.cw = (int)cw;
  a.ch = (int)ch;
  a.scale = L.scale;
  a.Q = *L.Q;

  const int esz = L.in_is_f32 ? 4 : 1;
  // 16-byte (float) / 4-byte (mono8) vector loads need aligned rows; a row tail may read up to 3 pixels past the
  // crop edge, which stays inside the frame row only if the border is at least 3 wide.
  const size_t valign = L.in_is_f32 ? 16 : 4;
  const bool vec = (reinterpret_cast<uintptr_t>(L.in) % valign == 0) && (L.step % valign == 0) &&
                   (L.frame_stride % valign == 0) && ((size_t)L.border * esz % valign == 0) &&
                   (cw % 4 == 0 || L.border >= 3) && !L.force_scalar;

  const int math = L.arith_fast ? kMathFast
                   : (a.Q.rectified && !L.force_generic ? (a.Q.q33_zero ? kMathRect0 : kMathRectW) : kMathGeneric);
  const bool compact = L.compact;

  int grid;
  if (compact) {
    const uint64_t pts = (uint64_t)cw * ch;
    a.tiles_per_frame = (uint32_t)((pts + kTilePts - 1) / kTilePts);
    a.tile_desc = static_cast<unsigned long long *>(L.scratch);
    a.ticket = L.ticket;
    a.epoch = L.epoch;
    const uint64_t total = (uint64_t)a.tiles_per_frame * L.n_frames;
    grid = (int)(total < (uint64_t)L.sm_count * 4 ? total : (uint64_t)L.sm_count * 4);
    cudaError_t e = cudaMemsetAsync(a.ticket, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
  } else {
    // rows per unit: large units amortise the per-unit constants, small ones spread a lone frame over the chip
    a.n_seg = (int)((cw + kSegCols - 1) / kSegCols);
    int rb = L.rows_per_unit > 0 ? L.rows_per_unit : 8;
    const uint64_t want_units = (uint64_t)L.sm_count * 32;
    while (rb > 2 && (uint64_t)a.n_seg * ((ch + rb - 1) / rb) * L.n_frames < want_units) rb >>= 1;
    if (rb > 32) rb = 32;
    rb &= ~1;
    if (rb < 2) rb = 2;
    a.rows_per_unit = rb;
    a.n_rb = (int)((ch + rb - 1) / rb);
    a.units_per_frame = (uint32_t)a.n_seg * (uint32_t)a.n_rb;
    const uint64_t total = (uint64_t)a.units_per_frame * L.n_frames;
    if (total > 0xffffffffull) return cudaErrorInvalidValue;
    a.total_units = (uint32_t)total;
    grid = (int)((total + kWarpsPerCta - 1) / kWarpsPerCta);
  }
  const int min_blocks = L.ctas_per_sm > 0 ? L.ctas_per_sm : 8;
  if (launches) *launches = 1;

#define D2PC_DISPATCH(T)                                                          \
  switch (math) {                                                                 \
    case kMathRect0: return launch_typed<T, kMathRect0>(a, vec, compact, grid, min_blocks, stream); \
    case kMathRectW: return launch_typed<T, kMathRectW>(a, vec, compact, grid, min_blocks, stream); \
    case kMathGeneric: return launch_typed<T, kMathGeneric>(a, vec, compact, grid, min_blocks, stream); \
    default: return launch_typed<T, kMathFast>(a, vec, compact, grid, min_blocks, stream);    \
  }
  if (L.in_is_f32) {
    D2PC_DISPATCH(float)
  } else {
    D2PC_DISPATCH(uint8_t)
  }
#undef D2PC_DISPATCH
}

}  // namespace d2pc
