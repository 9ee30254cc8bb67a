// experiments/median_variants.cu -- NOT compiled into libd2pc_b200.so, NOT shipped: three exact KxK median
// formulations that were built, tested bit-exact (every median test of tests/test_gpu_mono8.py passed with them as
// `median_variant` 3 / 4 / 5) and measured SLOWER than the window-histogram kernel of csrc/median.cu in round 2.
// Kept as source for the record; the measurements are in DESIGN.md section 2.4 and profiles/r2_median_*_ncu_full.txt.
// To revive one: paste it into median.cu (it uses MedianArgs + clampi from there, plus an `int src_aligned4` member
// for the SWAR-4 kernel) and dispatch on L.variant.
//
//   median_pair_kernel<K>        two adjacent outputs per thread, 16-bit histogram slots (12 instead of 22 updates
//                                per output at K = 11), ranks as 16-bit lanes via VIADDMNMX.S16x2.RELU in the slot
//                                offset domain.  256 x 752x480: 1.07 ms vs 0.91 ms; 504 M vs 603 M warp
//                                instructions (12 instructions per update where 7.5 were planned, two serial walks
//                                per thread), 10 warps per SM instead of 24.
//   median_swar_kernel<K, true>  four outputs per thread, 32-bit histogram words, shared-memory atomics: 1.69 ms --
//                                ATOMS retires one lane per clock per SM.
//   median_swar_kernel<K, false> the same with batched load / store updates: 1.47 ms -- 32 KB of histogram per warp
//                                leaves six warps per SM; the 624-instruction row body is latency-bound.
#if 0
// ---------------------------------------------------------------------------
// Pair histogram: two adjacent output columns per thread share 16-bit histogram slots
// ---------------------------------------------------------------------------
// The window-histogram kernel above is bound by the shared-memory pipe: 2K read-modify-writes per output and row,
// three LSU instructions each (ring entry, counter load, counter store).  The windows of two adjacent outputs
// share K - 1 of their K columns, so here a thread owns TWO adjacent outputs and a 16-bit slot per bin holds both
// 8-bit counters: a pixel of the K + 1 columns under the thread updates both windows with one 16-bit
// read-modify-write (increment 0x0001, 0x0101 or 0x0100 by column) -- 2(K + 1) updates per two outputs instead of
// 2K per output (K = 11: 12 instead of 22).  Layout: slot (bin, lane) at byte (bin >> 1) * 128 + 4 * lane +
// 2 * (bin & 1): lane L only ever touches bank L, whatever the data.  The two running ranks live in one register as
// 16-bit lanes and are maintained in the slot-offset domain (the offset is monotonic in the bin) with one packed
// DPX instruction per pixel: [off >= med_off] per lane = max(min(off + (1 - med_off), cap), 0), the column's
// membership in each window folded into cap.  Same sliding-down-a-strip structure, same ring of pre-transformed
// rows as the kernel above; 16 KB of histogram per warp (64 outputs): the same 256 B per output in flight.
constexpr int kPairWarps = 2;
constexpr int kPairThreads = kPairWarps * 32;
constexpr int kPairCols = 64;        // outputs per warp row
constexpr int kPairRingPitch = 80;   // >= 64 + 15 - 1 entries

__device__ __forceinline__ uint32_t pair_off(uint32_t bin) { return ((bin >> 1) << 7) | ((bin & 1u) << 1); }

template <int K>
__global__ void __launch_bounds__(kPairThreads) median_pair_kernel(const __grid_constant__ MedianArgs a) {
  constexpr int R = K / 2;
  constexpr int kRank = (K * K) / 2;
  constexpr int NCOL = kPairCols + K - 1;  // ring entries per row
  __shared__ __align__(16) uint8_t s_hist[kPairWarps][256 * 32 * 2];
  __shared__ __align__(4) uint16_t s_ring[kPairWarps][K][kPairRingPitch];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  uint8_t *hist = s_hist[wic];
  uint8_t *hb = hist + 4 * lane;  // this lane's slots: bin b at hb[pair_off(b)] (16 bits: low byte output 0, high byte output 1)
  uint16_t(*ring)[kPairRingPitch] = s_ring[wic];
  auto slot16 = [&](uint32_t off) -> uint16_t & { return *reinterpret_cast<uint16_t *>(hb + off); };
  // column dx (0..K) of the thread's K + 1 columns belongs to output 0 for dx < K and to output 1 for dx >= 1
  auto inc_of = [](int dx) { return (uint32_t)((dx < K ? 1 : 0) | (dx >= 1 ? 0x100 : 0)); };
  auto cap_of = [](int dx) { return (uint32_t)((dx < K ? 1 : 0) | (dx >= 1 ? 0x10000 : 0)); };

  for (uint32_t unit = blockIdx.x * kPairWarps + wic; unit < a.total_units; unit += gridDim.x * kPairWarps) {
    const uint32_t f = unit / a.units_per_frame;
    const uint32_t rem = unit - f * a.units_per_frame;
    const int strip = rem / a.n_colblk;
    const int cb = rem - strip * a.n_colblk;
    const int x0 = a.ox0 + cb * kPairCols;
    const int y_first = a.oy0 + strip * a.strip_rows;
    const int y_end = min(y_first + a.strip_rows, a.oy0 + a.oh);
    const uint8_t *src = a.src + (size_t)f * a.src_frame_stride;
    uint8_t *dst = a.dst + (size_t)f * a.dst_frame_stride;
    const int xo = x0 + 2 * lane;
    const int x_end = a.ox0 + a.ow;
    // columns this lane fetches for every ring row (replicate border = clamp): entries lane, 32 + lane, 64 + lane
    const int gx_a = clampi(x0 - R + lane, 0, a.width - 1);
    const int gx_b = clampi(x0 - R + 32 + lane, 0, a.width - 1);
    const int gx_c = clampi(x0 - R + 64 + lane, 0, a.width - 1);
    const bool has_c = 64 + lane < NCOL;

    // ---- zero the histogram (warp-cooperative, 16 B per store)
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < (256 * 32 * 2) / (32 * 16); ++i)
      reinterpret_cast<uint4 *>(hist)[i * 32 + lane] = make_uint4(0, 0, 0, 0);
    // ---- fill the ring with the window rows of the first output row
#pragma unroll
    for (int s = 0; s < K; ++s) {
      const uint8_t *row = src + (size_t)clampi(y_first - R + s, 0, a.height - 1) * a.src_step;
      ring[s][lane] = (uint16_t)pair_off(row[gx_a]);
      ring[s][32 + lane] = (uint16_t)pair_off(row[gx_b]);
      if (has_c) ring[s][64 + lane] = (uint16_t)pair_off(row[gx_c]);
    }
    __syncwarp();
#pragma unroll 1
    for (int s = 0; s < K; ++s) {
#pragma unroll
      for (int dx = 0; dx <= K; ++dx) {
        const uint32_t off = ring[s][2 * lane + dx];
        slot16(off) = (uint16_t)(slot16(off) + inc_of(dx));
      }
    }
    // ---- initial medians of the two outputs: two bins (one 32-bit word) at a time, then bin by bin
    int med[2], below[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int w = 0, bl = 0;
      for (; w < 128; ++w) {
        const uint32_t word = *reinterpret_cast<const uint32_t *>(hb + (w << 7));
        const int s2 = (int)((word >> (8 * i)) & 0xff) + (int)((word >> (16 + 8 * i)) & 0xff);
        if (bl + s2 > kRank) break;
        bl += s2;
      }
      int m = 2 * w;
      for (;;) {
        const int hm = hb[pair_off(m) + i];
        if (bl + hm > kRank) break;
        bl += hm;
        ++m;
      }
      med[i] = m, below[i] = bl;
    }

    int slot = 0;  // ring slot holding the oldest window row
    uint32_t na = 0, nb = 0, nc = 0;
    if (y_first + 1 < y_end) {
      const uint8_t *row = src + (size_t)clampi(y_first + 1 + R, 0, a.height - 1) * a.src_step;
      na = row[gx_a], nb = row[gx_b], nc = has_c ? row[gx_c] : 0u;
    }
    for (int y = y_first; y < y_end; ++y) {
      {
        uint8_t *o = dst + (size_t)y * a.dst_step + xo;
        if (xo + 1 < x_end && (reinterpret_cast<uintptr_t>(o) & 1u) == 0) {
          *reinterpret_cast<uint16_t *>(o) = (uint16_t)(med[0] | (med[1] << 8));
        } else {
          if (xo < x_end) o[0] = (uint8_t)med[0];
          if (xo + 1 < x_end) o[1] = (uint8_t)med[1];
        }
      }
      if (y + 1 >= y_end) break;
      // the row entering the window was fetched one iteration ago; fetch the one after it now
      const uint32_t ea = pair_off(na), eb = pair_off(nb), ec = pair_off(nc);
      if (y + 2 < y_end) {
        const uint8_t *row = src + (size_t)clampi(y + 2 + R, 0, a.height - 1) * a.src_step;
        na = row[gx_a], nb = row[gx_b], nc = has_c ? row[gx_c] : 0u;
      }
      const uint32_t mo0 = pair_off(med[0]), mo1 = pair_off(med[1]);
      const uint32_t t = ((1u - mo0) & 0xffffu) | ((1u - mo1) << 16);  // 1 - med_off per 16-bit lane
      uint32_t bw = (uint32_t)below[0] | ((uint32_t)below[1] << 16);   // the two ranks as 16-bit lanes
      // remove the oldest row (entries read up front: a ring load can not move across a histogram store)
      uint32_t offs[K + 1];
#pragma unroll
      for (int dx = 0; dx <= K; ++dx) offs[dx] = ring[slot][2 * lane + dx];
#pragma unroll
      for (int dx = 0; dx <= K; ++dx) {
        slot16(offs[dx]) = (uint16_t)(slot16(offs[dx]) - inc_of(dx));
        bw += __viaddmin_s16x2_relu(offs[dx] * 0x10001u, t, cap_of(dx));  // a pixel >= med leaves: below is unchanged, but
      }                                                                    // the caps cancel against the additions below
      __syncwarp();
      ring[slot][lane] = (uint16_t)ea;
      ring[slot][32 + lane] = (uint16_t)eb;
      if (has_c) ring[slot][64 + lane] = (uint16_t)ec;
      __syncwarp();
#pragma unroll
      for (int dx = 0; dx <= K; ++dx) offs[dx] = ring[slot][2 * lane + dx];
#pragma unroll
      for (int dx = 0; dx <= K; ++dx) {
        slot16(offs[dx]) = (uint16_t)(slot16(offs[dx]) + inc_of(dx));
        bw -= __viaddmin_s16x2_relu(offs[dx] * 0x10001u, t, cap_of(dx));
      }
      slot = (slot + 1 == K) ? 0 : slot + 1;
      below[0] = (int)(bw & 0xffffu), below[1] = (int)(bw >> 16);
      // re-centre both outputs: invariant below <= kRank < below + hist[med]
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int bl = below[i], m = med[i];
        while (bl > kRank) {
          --m;
          bl -= hb[pair_off(m) + i];
        }
        for (;;) {
          const int hm = hb[pair_off(m) + i];
          if (bl + hm > kRank) break;
          bl += hm;
          ++m;
        }
        below[i] = bl, med[i] = m;
      }
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

template <int K>
cudaError_t launch_swar(const MedianArgs &a, int grid, bool atomic, cudaStream_t s) {
  if (atomic) median_swar_kernel<K, true><<<grid, 32, 0, s>>>(a);
  else median_swar_kernel<K, false><<<grid, 32, 0, s>>>(a);
  return cudaGetLastError();
}

template <int K>
cudaError_t launch_pair(const MedianArgs &a, int grid, cudaStream_t s) {
  median_pair_kernel<K><<<grid, kPairThreads, 0, s>>>(a);
  return cudaGetLastError();
}

#endif
