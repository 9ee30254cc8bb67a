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
  int src_aligned4;  // src base and row step are multiples of 4 (word fetches allowed)
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
// SWAR-4 sliding histogram: four adjacent output columns per thread share one histogram word per bin
// ---------------------------------------------------------------------------
// The window histogram kernel above pays 2K shared-memory read-modify-writes per output.  The windows of
// horizontally adjacent outputs overlap in all but one column, so here a thread owns FOUR adjacent outputs and
// keeps their four 8-bit counters of a bin in one 32-bit word: a pixel that enters (or leaves) the K + 3 columns
// under the thread updates the counters of every output whose window holds that column with ONE 32-bit add
// (increment = one 0x01 byte per affected output; counters never exceed K*K <= 225, so bytes never carry).
// That is 2(K+3) updates per 4 outputs instead of 2K per output (K = 11: 7 instead of 22), each of them a
// fire-and-forget shared-memory atomic (ATOMS.ADD, no result, so nothing serialises on a load->add->store chain
// and two pixels of one row that fall into the same bin need no special care).
// The running rank (#window pixels below the current median) of the four outputs lives in two registers as
// 16-bit lanes and is maintained with the packed DPX instruction VIADDMNMX.S16x2.RELU:
//   [p >= med] per lane = max(min(p + (1 - med), c), 0), c = 1 where the column belongs to that output's window.
// A warp owns 128 output columns x a strip of rows; layout hist[bin][lane] (word), so lane L only touches bank L.
constexpr int kSwarOut = 4;                 // outputs per thread
constexpr int kSwarCols = 32 * kSwarOut;    // outputs per warp row

template <int K>
struct Swar {
  static constexpr int R = K / 2;
  static constexpr int NC = K + kSwarOut - 1;  // image columns under one thread
  static constexpr int NW = (NC + 3) / 4;      // ring words a thread reads per row
  static constexpr int RW = 32 + NW;           // ring words per row (word j = columns x0 - R + 4j .. + 3)
  // outputs i (0..3) whose window holds thread-local column c: c - 2R <= i <= c
  static constexpr __host__ __device__ bool has(int c, int i) { return i <= c && i >= c - 2 * R; }
  static constexpr __host__ __device__ uint32_t inc(int c) {
    return (has(c, 0) ? 1u : 0u) | (has(c, 1) ? 1u << 8 : 0u) | (has(c, 2) ? 1u << 16 : 0u) | (has(c, 3) ? 1u << 24 : 0u);
  }
  static constexpr __host__ __device__ uint32_t cap01(int c) { return (has(c, 0) ? 1u : 0u) | (has(c, 1) ? 1u << 16 : 0u); }
  static constexpr __host__ __device__ uint32_t cap23(int c) { return (has(c, 2) ? 1u : 0u) | (has(c, 3) ? 1u << 16 : 0u); }
};

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return (w >> (8 * i)) & 0xffu; }

template <int K, bool kAtomic>
__global__ void __launch_bounds__(32) median_swar_kernel(const __grid_constant__ MedianArgs a) {
  using S = Swar<K>;
  constexpr int R = S::R, NC = S::NC, NW = S::NW, RW = S::RW;
  constexpr int kRank = (K * K) / 2;
  __shared__ __align__(16) uint32_t hist[256 * 32];
  __shared__ __align__(16) uint32_t ring[K * RW];
  const int lane = threadIdx.x;
  uint32_t *hl = hist + lane;  // counters of bin b for this lane's four outputs: hl[b * 32]

  for (uint32_t unit = blockIdx.x; unit < a.total_units; unit += gridDim.x) {
    const uint32_t f = unit / a.units_per_frame;
    const uint32_t rem = unit - f * a.units_per_frame;
    const int strip = rem / a.n_colblk;
    const int cb = rem - strip * a.n_colblk;
    const int x0 = a.ox0 + cb * kSwarCols;
    const int y_first = a.oy0 + strip * a.strip_rows;
    const int y_end = min(y_first + a.strip_rows, a.oy0 + a.oh);
    const uint8_t *src = a.src + (size_t)f * a.src_frame_stride;
    uint8_t *dst = a.dst + (size_t)f * a.dst_frame_stride;
    const int xo = x0 + kSwarOut * lane;    // first of this thread's four output columns
    const int x_end = a.ox0 + a.ow;
    // Ring rows are fetched as aligned 32-bit words when the whole footprint (plus one word) lies inside the
    // image and rows are word aligned; otherwise byte by byte with the replicate border (clamp).
    const int c0 = x0 - R;  // image column of ring byte 0
    const bool fast = a.src_aligned4 && c0 >= 0 && c0 + 4 * RW + 4 <= a.width;

    // A row fetch is split in two so the global loads of row y + 1 are in flight while row y is processed:
    // fetch_issue() only loads (raw words / bytes), fetch_finish() aligns / packs them into the ring words.
    auto fetch_issue = [&](int y, uint32_t (&r)[8]) {
      const uint8_t *row = src + (size_t)min(max(y, 0), a.height - 1) * a.src_step;
      if (fast) {
        const uint32_t *gw = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(row + c0) & ~(uintptr_t)3);
        r[0] = __ldg(gw + lane), r[1] = __ldg(gw + lane + 1);
        if (lane < NW) r[2] = __ldg(gw + 32 + lane), r[3] = __ldg(gw + 33 + lane);
      } else {
        auto px = [&](int c) { return (uint32_t)row[min(max(c, 0), a.width - 1)]; };
        const int c = c0 + 4 * lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = px(c + k);
        if (lane < NW) {
#pragma unroll
          for (int k = 0; k < 4; ++k) r[4 + k] = px(c + 128 + k);
        }
      }
    };
    auto fetch_finish = [&](const uint32_t (&r)[8], uint32_t &w0, uint32_t &w1) {
      if (fast) {
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src + c0) & 3u) * 8u;  // rows are word aligned
        w0 = __funnelshift_r(r[0], r[1], sh);
        w1 = __funnelshift_r(r[2], r[3], sh);
      } else {
        w0 = r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24);
        w1 = r[4] | (r[5] << 8) | (r[6] << 16) | (r[7] << 24);
      }
    };
    auto fetch_row = [&](int y, uint32_t &w0, uint32_t &w1) {
      uint32_t r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      fetch_issue(y, r);
      fetch_finish(r, w0, w1);
    };

    // ---- zero the histogram, load the K window rows of the first output row
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < (256 * 32) / (32 * 4); ++i) reinterpret_cast<uint4 *>(hist)[i * 32 + lane] = make_uint4(0, 0, 0, 0);
#pragma unroll 1
    for (int s = 0; s < K; ++s) {
      uint32_t w0, w1;
      fetch_row(y_first - R + s, w0, w1);
      ring[s * RW + lane] = w0;
      if (lane < NW) ring[s * RW + 32 + lane] = w1;
    }
    __syncwarp();
#pragma unroll 1
    for (int s = 0; s < K; ++s) {
      uint32_t w[NW];
#pragma unroll
      for (int j = 0; j < NW; ++j) w[j] = ring[s * RW + lane + j];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const uint32_t p = byte_of(w[c >> 2], c & 3);
        if (kAtomic) atomicAdd(&hl[p * 32], S::inc(c));
        else hl[p * 32] += S::inc(c);
      }
    }
    // ---- initial medians of the four outputs at once: inclusive prefix per byte, med = #bins whose prefix <= rank
    int med[kSwarOut], below[kSwarOut];
    {
      uint32_t acc = 0, cnt = 0, bel = 0;
      for (int bin0 = 0; bin0 < 256; bin0 += 8) {
        uint32_t h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) h[k] = hl[(bin0 + k) * 32];
        uint32_t t = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          acc += h[k];
          t = acc + (uint32_t)(0x7f - kRank) * 0x01010101u;  // bit 7 of a byte: prefix > rank
          const uint32_t le = (~t >> 7) & 0x01010101u;
          cnt += le;
          bel += h[k] & (le * 0xffu);
        }
        if ((t & 0x80808080u) == 0x80808080u) break;
      }
#pragma unroll
      for (int i = 0; i < kSwarOut; ++i) med[i] = (int)byte_of(cnt, i), below[i] = (int)byte_of(bel, i);
    }

    int slot = 0;  // ring slot holding the oldest window row
    uint32_t raw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (y_first + 1 < y_end) fetch_issue(y_first + 1 + R, raw);
    for (int y = y_first; y < y_end; ++y) {
      {  // store the four medians
        uint8_t *o = dst + (size_t)y * a.dst_step + xo;
        const uint32_t packed = (uint32_t)med[0] | ((uint32_t)med[1] << 8) | ((uint32_t)med[2] << 16) | ((uint32_t)med[3] << 24);
        if (xo + 3 < x_end && (reinterpret_cast<uintptr_t>(o) & 3u) == 0) {
          *reinterpret_cast<uint32_t *>(o) = packed;
        } else {
#pragma unroll
          for (int i = 0; i < kSwarOut; ++i)
            if (xo + i < x_end) o[i] = (uint8_t)med[i];
        }
      }
      if (y + 1 >= y_end) break;
      uint32_t n0, n1;
      fetch_finish(raw, n0, n1);                       // row y + 1 + R, loaded one iteration ago
      if (y + 2 < y_end) fetch_issue(y + 2 + R, raw);  // row y + 2 + R: consumed in the next iteration
      uint32_t ow[NW], nw[NW];
#pragma unroll
      for (int j = 0; j < NW; ++j) ow[j] = ring[slot * RW + lane + j];
      __syncwarp();
      ring[slot * RW + lane] = n0;
      if (lane < NW) ring[slot * RW + 32 + lane] = n1;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NW; ++j) nw[j] = ring[slot * RW + lane + j];
      slot = (slot + 1 == K) ? 0 : slot + 1;

      // rank words: below counts as 16-bit lanes; t = 1 - med per lane
      uint32_t b01 = (uint32_t)below[0] | ((uint32_t)below[1] << 16), b23 = (uint32_t)below[2] | ((uint32_t)below[3] << 16);
      const uint32_t t01 = ((uint32_t)(1 - med[0]) & 0xffffu) | ((uint32_t)(1 - med[1]) << 16);
      const uint32_t t23 = ((uint32_t)(1 - med[2]) & 0xffffu) | ((uint32_t)(1 - med[3]) << 16);
      uint32_t po[NC], pn[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        po[c] = __byte_perm(ow[c >> 2], 0, 0x4040 | (c & 3) | ((c & 3) << 8));  // p | p << 16
        pn[c] = __byte_perm(nw[c >> 2], 0, 0x4040 | (c & 3) | ((c & 3) << 8));
        if (S::cap01(c)) b01 = b01 + __viaddmin_s16x2_relu(po[c], t01, S::cap01(c)) - __viaddmin_s16x2_relu(pn[c], t01, S::cap01(c));
        if (S::cap23(c)) b23 = b23 + __viaddmin_s16x2_relu(po[c], t23, S::cap23(c)) - __viaddmin_s16x2_relu(pn[c], t23, S::cap23(c));
      }
      if (kAtomic) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          atomicAdd(hl + ((po[c] & 0xffu) << 5), 0u - S::inc(c));
          atomicAdd(hl + ((pn[c] & 0xffu) << 5), S::inc(c));
        }
      } else {
        // Plain loads / stores, kBatch columns (2 * kBatch counters) at a time: all loads of a batch are issued
        // before its stores, and every counter is stored as (loaded value + the deltas of ALL batch members that
        // address the same word), so members that coincide (a pixel leaving and one entering the same bin, equal
        // neighbours) each store the same, complete value and their order does not matter.
        constexpr int kBatch = 2;
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += kBatch) {
          constexpr int M = 2 * kBatch;
          uint32_t off[M], val[M], tot[M], dlt[M];
          bool live[M];
#pragma unroll
          for (int j = 0; j < M; ++j) {
            const int c = c0 + (j >> 1);
            live[j] = c < NC;
            const uint32_t p = live[j] ? ((j & 1) ? pn[c] : po[c]) : 0u;
            off[j] = (p & 0xffu) << 5;
            dlt[j] = live[j] ? ((j & 1) ? S::inc(c < NC ? c : 0) : 0u - S::inc(c < NC ? c : 0)) : 0u;
          }
#pragma unroll
          for (int j = 0; j < M; ++j)
            if (live[j]) val[j] = hl[off[j]];
#pragma unroll
          for (int j = 0; j < M; ++j) {
            tot[j] = dlt[j];
#pragma unroll
            for (int i = 0; i < M; ++i)
              if (i != j && live[i] && live[j]) tot[j] += (off[i] == off[j]) ? dlt[i] : 0u;
          }
#pragma unroll
          for (int j = 0; j < M; ++j)
            if (live[j]) hl[off[j]] = val[j] + tot[j];
        }
      }
      below[0] = (int)(b01 & 0xffffu), below[1] = (int)(b01 >> 16), below[2] = (int)(b23 & 0xffffu), below[3] = (int)(b23 >> 16);
      // Re-centre (invariant below <= rank < below + hist[med]).  First one look per output with the four loads in
      // flight together -- that settles an output whose median moved by at most one bin -- then a tight loop per
      // output for the ones that have further to go (a window crossing a depth edge).
      const uint8_t *hb = reinterpret_cast<const uint8_t *>(hl);
      bool ok[kSwarOut];
      {
        int v[kSwarOut], mm[kSwarOut];
#pragma unroll
        for (int i = 0; i < kSwarOut; ++i) {
          mm[i] = med[i] - (below[i] > kRank ? 1 : 0);
          v[i] = (int)hb[mm[i] * 128 + i];
        }
#pragma unroll
        for (int i = 0; i < kSwarOut; ++i) {
          if (below[i] > kRank) {
            med[i] = mm[i];
            below[i] -= v[i];
            ok[i] = below[i] <= kRank;
          } else if (below[i] + v[i] <= kRank) {
            below[i] += v[i];
            med[i] = mm[i] + 1;
            ok[i] = false;
          } else {
            ok[i] = true;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < kSwarOut; ++i) {
        if (ok[i]) continue;
        int bl = below[i], m = med[i];
        if (bl > kRank) {
          do {
            --m;
            bl -= (int)hb[m * 128 + i];
          } while (bl > kRank);
        } else {
          for (;;) {
            const int hm = (int)hb[m * 128 + i];
            if (bl + hm > kRank) break;
            bl += hm;
            ++m;
          }
        }
        below[i] = bl, med[i] = m;
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
cudaError_t launch_swar(const MedianArgs &a, int grid, bool atomic, cudaStream_t s) {
  if (atomic) median_swar_kernel<K, true><<<grid, 32, 0, s>>>(a);
  else median_swar_kernel<K, false><<<grid, 32, 0, s>>>(a);
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
  a.src_aligned4 = (reinterpret_cast<uintptr_t>(L.src) % 4 == 0 && L.src_step % 4 == 0 && L.src_frame_stride % 4 == 0) ? 1 : 0;
  if (launches) *launches = 1;
  if (L.ksize == 3 && L.variant == 0 && L.oh <= 65535 && L.n_frames <= 65535) {
    median3_net_kernel<<<dim3((unsigned)((L.ow + 127) / 128), (unsigned)L.oh, L.n_frames), 128, 0, stream>>>(a);
    return cudaGetLastError();
  }
  // variant 0 (default) and 2: one window histogram per output column; 4: SWAR-4 histogram, load/store updates;
  // 3: SWAR-4 with shared-memory atomics (both measured slower)
  const bool swar = L.variant == 3 || L.variant == 4;
  const int cols_per_unit = swar ? kSwarCols : 32;
  a.n_colblk = (L.ow + cols_per_unit - 1) / cols_per_unit;
  // strip height: long strips amortise the start-up of a unit (zeroing + K rows of updates + the first rank
  // search: ~9 rows' worth for the SWAR kernel), short ones fill the chip with warps when there is one frame.
  const int min_strip = swar ? 4 : 2;
  const uint64_t want = (uint64_t)L.sm_count * (swar ? 6 : 16);
  int strip = 64;
  while (strip > min_strip && (uint64_t)a.n_colblk * ((L.oh + strip - 1) / strip) * L.n_frames < want) strip >>= 1;
  if (L.strip_rows > 0) strip = L.strip_rows;
  a.strip_rows = strip;
  a.n_strip = (L.oh + strip - 1) / strip;
  a.units_per_frame = (uint32_t)a.n_colblk * (uint32_t)a.n_strip;
  const uint64_t total = (uint64_t)a.units_per_frame * L.n_frames;
  if (total > 0xffffffffull) return cudaErrorInvalidValue;
  a.total_units = (uint32_t)total;
  const uint64_t cap = (uint64_t)L.sm_count * 6;
  if (swar) {
    const int grid = (int)(total < cap ? total : cap);  // one warp per CTA, 34 KB of shared memory each
    const bool atomic = L.variant == 3;
    switch (L.ksize) {
      case 3: return launch_swar<3>(a, grid, atomic, stream);
      case 5: return launch_swar<5>(a, grid, atomic, stream);
      case 7: return launch_swar<7>(a, grid, atomic, stream);
      case 9: return launch_swar<9>(a, grid, atomic, stream);
      case 11: return launch_swar<11>(a, grid, atomic, stream);
      case 13: return launch_swar<13>(a, grid, atomic, stream);
      default: return launch_swar<15>(a, grid, atomic, stream);
    }
  }
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
