/*
 * d2pc_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See d2pc_oracle.h for scope, the parity pin and who may call this.
 * Build with -ffp-contract=off: every rounding below is deliberate.
 * All file:line citations are relative to /root/reference.
 */
#include "d2pc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* Q from intrinsics                                                          */
/* ------------------------------------------------------------------------- */

/* disparity_to_point_cloud.hpp:90-104 calls cv::stereoRectify with zero
 * distortion, R = I and t = (-b, 0, 0).  For that call OpenCV (calib3d,
 * stereoRectify) reduces to: both rectifying rotations are I; the new focal
 * length is K[1][1] (the "other axis" focal length of a horizontal rig); the
 * new principal point is (n-1)/2 minus the mean of the four image corners
 * pushed through undistortPoints (float32 result, multiply by 1/f) and
 * projectPoints with the new focal length and a zero principal point
 * (float32 result); both cameras share K so the ZERO_DISPARITY averaging is a
 * no-op; alpha = -1 skips the scaling step.  Then
 *   Q = [1 0 0 -cx'; 0 1 0 -cy'; 0 0 0 f'; 0 0 -1/tx (cx1'-cx2')/tx].
 * Checked bit-for-bit against cv2.stereoRectify (tests/golden). */
int d2pc_oracle_q_from_intrinsics(double fx, double fy, double cx, double cy,
                                  double baseline, int rect_w, int rect_h,
                                  double q[16]) {
  if (!(fx != 0.0) || !(fy != 0.0) || !(baseline != 0.0) || rect_w <= 0 ||
      rect_h <= 0)
    return -1;
  const double f_new = fy;
  const double ifx = 1.0 / fx, ify = 1.0 / fy;
  double sum_x = 0.0, sum_y = 0.0;
  for (int i = 0; i < 4; ++i) {
    const int j = i < 2 ? 0 : 1;
    const float px = (float)((i % 2) * (rect_w - 1));
    const float py = (float)(j * (rect_h - 1));
    const float nx = (float)(((double)px - cx) * ifx);
    const float ny = (float)(((double)py - cy) * ify);
    const float qx = (float)((double)nx * f_new + 0.0);
    const float qy = (float)((double)ny * f_new + 0.0);
    sum_x += (double)qx;
    sum_y += (double)qy;
  }
  const double cc_x = (rect_w - 1) * 0.5 - sum_x / 4.0;
  const double cc_y = (rect_h - 1) * 0.5 - sum_y / 4.0;
  const double tx = -baseline;
  memset(q, 0, 16 * sizeof(double));
  q[0] = 1.0;
  q[3] = -cc_x;
  q[5] = 1.0;
  q[7] = -cc_y;
  q[11] = f_new;
  q[14] = -1.0 / tx;
  q[15] = (cc_x - cc_x) / tx;
  return 0;
}

/* ------------------------------------------------------------------------- */
/* medianBlur, CV_8UC1, replicate border                                      */
/* ------------------------------------------------------------------------- */

static inline int clampi(int v, int lo, int hi) {
  return v < lo ? lo : (v > hi ? hi : v);
}

/* Sliding-histogram (Huang) median with a 16-bin coarse level.  The result is
 * the exact rank-(k*k/2) order statistic, which is what every OpenCV code
 * path (sorting network, O(m), O(1)) returns for 8-bit input. */
void d2pc_oracle_median_blur_u8(const uint8_t *src, int w, int h,
                                size_t src_step, uint8_t *dst, size_t dst_step,
                                int ksize) {
  const int r = ksize / 2;
  const int rank = (ksize * ksize) / 2; /* 0-based index of the median */
  const uint8_t **rows = (const uint8_t **)malloc(sizeof(*rows) * ksize);
  for (int y = 0; y < h; ++y) {
    int fine[256];
    int coarse[16];
    memset(fine, 0, sizeof fine);
    memset(coarse, 0, sizeof coarse);
    for (int dy = -r; dy <= r; ++dy)
      rows[dy + r] = src + (size_t)clampi(y + dy, 0, h - 1) * src_step;
    for (int dy = 0; dy < ksize; ++dy)
      for (int dx = -r; dx <= r; ++dx) {
        const uint8_t v = rows[dy][clampi(dx, 0, w - 1)];
        ++fine[v];
        ++coarse[v >> 4];
      }
    for (int x = 0; x < w; ++x) {
      if (x > 0) {
        const int xo = clampi(x - r - 1, 0, w - 1);
        const int xn = clampi(x + r, 0, w - 1);
        for (int dy = 0; dy < ksize; ++dy) {
          const uint8_t vo = rows[dy][xo], vn = rows[dy][xn];
          --fine[vo];
          --coarse[vo >> 4];
          ++fine[vn];
          ++coarse[vn >> 4];
        }
      }
      int acc = 0, c = 0;
      while (acc + coarse[c] <= rank) acc += coarse[c++];
      int b = c << 4;
      while (acc + fine[b] <= rank) acc += fine[b++];
      dst[(size_t)y * dst_step + x] = (uint8_t)b;
    }
  }
  free(rows);
}

/* ------------------------------------------------------------------------- */
/* convertTo                                                                  */
/* ------------------------------------------------------------------------- */

/* disparity_to_point_cloud.cpp:61.  OpenCV's 8u->32f scaled convert works in
 * float: dst = (float)src * (float)alpha + 0.f.  alpha = 1/8 is exact. */
void d2pc_oracle_convert_u8_f32(const uint8_t *src, int w, int h,
                                size_t src_step, float *dst, size_t dst_step,
                                double alpha) {
  const float a = (float)alpha;
  for (int y = 0; y < h; ++y) {
    const uint8_t *s = src + (size_t)y * src_step;
    float *d = (float *)((uint8_t *)dst + (size_t)y * dst_step);
    for (int x = 0; x < w; ++x) d[x] = (float)s[x] * a + 0.0f;
  }
}

/* ------------------------------------------------------------------------- */
/* reprojectImageTo3D                                                         */
/* ------------------------------------------------------------------------- */

/* disparity_to_point_cloud.cpp:64.  Rounding sequence pinned against
 * cv2 4.13.0: homogeneous vector in float64 in this exact
 * association, X/Y/Z cast to float32, divided by float64 W, cast to float32. */
void d2pc_oracle_reproject_image_to_3d(const float *disp, int w, int h,
                                       size_t disp_step, const double q[16],
                                       float *xyz) {
  for (int v = 0; v < h; ++v) {
    const float *drow = (const float *)((const uint8_t *)disp +
                                        (size_t)v * disp_step);
    float *o = xyz + (size_t)v * w * 3;
    const double dv = (double)v;
    for (int u = 0; u < w; ++u) {
      const double du = (double)u;
      const double d = (double)drow[u];
      double hh[4];
      for (int i = 0; i < 4; ++i)
        hh[i] = ((q[4 * i + 0] * du + q[4 * i + 1] * dv) + q[4 * i + 2] * d) +
                q[4 * i + 3] * 1.0;
      const float xf = (float)hh[0], yf = (float)hh[1], zf = (float)hh[2];
      o[3 * u + 0] = (float)((double)xf / hh[3]);
      o[3 * u + 1] = (float)((double)yf / hh[3]);
      o[3 * u + 2] = (float)((double)zf / hh[3]);
    }
  }
}

/* ------------------------------------------------------------------------- */
/* crop + pack                                                                */
/* ------------------------------------------------------------------------- */

typedef struct {
  float x, y, z, pad; /* pcl::PointXYZ: 16 bytes, 4th float is 1.0f */
} oracle_point;

/* disparity_to_point_cloud.cpp:69-85.  The push_back loop runs on an
 * un-reserved std::vector (capacity doubles, elements move on regrowth), then
 * pcl::toROSMsg memcpy's the whole array into PointCloud2.data. */
size_t d2pc_oracle_crop_pack(const float *xyz, int w, int h, int border,
                             uint8_t *cloud) {
  const long cw = (long)w - 2L * border, ch = (long)h - 2L * border;
  const size_t n = (cw > 0 && ch > 0) ? (size_t)cw * (size_t)ch : 0;
  if (!cloud || n == 0) return n;
  oracle_point *vec = NULL;
  size_t size = 0, cap = 0;
  for (int v = border; v < h - border; ++v) {
    const float *pv = xyz + (size_t)v * w * 3;
    for (int u = border; u < w - border; ++u) {
      if (size == cap) {
        const size_t ncap = cap ? 2 * cap : 1;
        oracle_point *nv =
            (oracle_point *)aligned_alloc(16, ncap * sizeof(oracle_point));
        if (size) memcpy(nv, vec, size * sizeof(oracle_point));
        free(vec);
        vec = nv;
        cap = ncap;
      }
      oracle_point p = {pv[3 * u + 0], pv[3 * u + 1], pv[3 * u + 2], 1.0f};
      vec[size++] = p;
    }
  }
  memcpy(cloud, vec, size * sizeof(oracle_point));
  free(vec);
  return size;
}

size_t d2pc_oracle_disparity_cb_f32(const float *disp, int w, int h,
                                    size_t step, const double q[16],
                                    uint8_t *cloud) {
  if (w <= 0 || h <= 0) return 0;
  float *xyz = (float *)malloc((size_t)w * h * 3 * sizeof(float));
  d2pc_oracle_reproject_image_to_3d(disp, w, h, step, q, xyz);
  const size_t n = d2pc_oracle_crop_pack(xyz, w, h, 40, cloud);
  free(xyz);
  return n;
}

size_t d2pc_oracle_disparity_cb_mono8(const uint8_t *img, int w, int h,
                                      size_t step, const double q[16],
                                      uint8_t *cloud) {
  if (w <= 0 || h <= 0) return 0;
  /* :50 cv_bridge::toCvCopy -> dense deep copy */
  uint8_t *copy = (uint8_t *)malloc((size_t)w * h);
  for (int y = 0; y < h; ++y)
    memcpy(copy + (size_t)y * w, img + (size_t)y * step, (size_t)w);
  /* :55-57 */
  uint8_t *med = (uint8_t *)malloc((size_t)w * h);
  d2pc_oracle_median_blur_u8(copy, w, h, (size_t)w, med, (size_t)w, 11);
  /* :60-61 */
  float *real = (float *)malloc((size_t)w * h * sizeof(float));
  d2pc_oracle_convert_u8_f32(med, w, h, (size_t)w, real, (size_t)w * 4,
                             1.0 / 8.0);
  /* :63-85 */
  const size_t n =
      d2pc_oracle_disparity_cb_f32(real, w, h, (size_t)w * 4, q, cloud);
  free(real);
  free(med);
  free(copy);
  return n;
}

size_t d2pc_oracle_filter_finite(const uint8_t *cloud, size_t n_points,
                                 uint8_t *out) {
  size_t k = 0;
  for (size_t i = 0; i < n_points; ++i) {
    oracle_point p;
    memcpy(&p, cloud + 16 * i, 16);
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      if (out) memcpy(out + 16 * k, &p, 16);
      ++k;
    }
  }
  return k;
}

/* ------------------------------------------------------------------------- */
/* ROS1 wire format of the published PointCloud2                              */
/* ------------------------------------------------------------------------- */

typedef struct {
  uint8_t *p;
  size_t cap, len;
} wbuf;

static void put(wbuf *b, const void *src, size_t n) {
  if (b->p && b->len + n <= b->cap) memcpy(b->p + b->len, src, n);
  b->len += n;
}
static void put_u32(wbuf *b, uint32_t v) { put(b, &v, 4); } /* little endian host */
static void put_u8(wbuf *b, uint8_t v) { put(b, &v, 1); }
static void put_str(wbuf *b, const char *s) {
  const uint32_t n = (uint32_t)strlen(s);
  put_u32(b, n);
  put(b, s, n);
}

/* disparity_to_point_cloud.cpp:79-90 + pcl::toROSMsg<PointXYZ> field table. */
size_t d2pc_oracle_serialize_pointcloud2(uint32_t seq, uint32_t sec,
                                         uint32_t nsec, const char *frame_id,
                                         const uint8_t *points,
                                         uint32_t n_points, uint8_t is_dense,
                                         uint8_t *out, size_t cap) {
  wbuf b = {out, cap, 0};
  static const char *names[3] = {"x", "y", "z"};
  put_u32(&b, seq);
  put_u32(&b, sec);
  put_u32(&b, nsec);
  put_str(&b, frame_id);
  put_u32(&b, 1);        /* height (:80) */
  put_u32(&b, n_points); /* width  (:79) */
  put_u32(&b, 3);
  for (uint32_t i = 0; i < 3; ++i) {
    put_str(&b, names[i]);
    put_u32(&b, 4 * i); /* offset */
    put_u8(&b, 7);      /* FLOAT32 */
    put_u32(&b, 1);     /* count */
  }
  put_u8(&b, 0);               /* is_bigendian */
  put_u32(&b, 16);             /* point_step = sizeof(pcl::PointXYZ) */
  put_u32(&b, 16u * n_points); /* row_step */
  put_u32(&b, 16u * n_points); /* data length */
  put(&b, points, (size_t)16 * n_points);
  put_u8(&b, is_dense); /* :81 false */
  return b.len;
}

/* ------------------------------------------------------------------------- */
/* depth_map_fusion                                                           */
/* ------------------------------------------------------------------------- */

static inline int absi(int v) { return v < 0 ? -v : v; }
static inline int maxi(int a, int b) { return a > b ? a : b; }
static inline int mini(int a, int b) { return a < b ? a : b; }

/* depth_map_fusion.cpp:247-265 */
int d2pc_oracle_crop_to_square(int cols, int rows, int offset_x, int offset_y,
                               int member_offset_y, int rect[4]) {
  const int num_cols = cols - absi(offset_x);
  const int num_rows = rows - absi(offset_y);
  const int n = mini(cols, rows) - maxi(absi(offset_x), absi(member_offset_y));
  int start_col, start_row;
  if (num_cols < num_rows) {
    start_col = maxi(0, offset_x);
    start_row = maxi(0, offset_y + (num_rows - num_cols) / 2);
  } else {
    start_col = maxi(0, offset_x + (num_cols - num_rows) / 2);
    start_row = maxi(0, offset_y);
  }
  rect[0] = start_col;
  rect[1] = start_row;
  rect[2] = n;
  rect[3] = n;
  if (n < 0 || start_col + n > cols || start_row + n > rows) return -1;
  return 0;
}

/* depth_map_fusion.cpp:268-273: transpose then flip around the vertical
 * axis, i.e. rot(r, c) = src(h-1-c, r); rot has h columns and w rows. */
void d2pc_oracle_rotate_cw(const uint8_t *src, int w, int h, size_t src_step,
                           uint8_t *dst) {
  for (int r = 0; r < w; ++r)
    for (int c = 0; c < h; ++c)
      dst[(size_t)r * h + c] = src[(size_t)(h - 1 - c) * src_step + r];
}

/* depth_map_fusion.cpp:219-235 */
int d2pc_oracle_grad_filter(int dist1, int dist2, int score1, int score2,
                            int grad1, int grad2) {
  (void)grad1;
  (void)grad2;
  const int thres = 100;
  const int tooClose = 230;
  const float relative_diff = (float)dist1 / (float)dist2;
  if (score1 < score2 && score1 < thres && dist1 < tooClose) {
    return dist1;
  } else if (score2 < score1 && score2 < thres && dist2 < tooClose) {
    return dist2;
  } else if (0.8 < (double)relative_diff && (double)relative_diff < 1.25 &&
             (double)score1 < 1.25 * thres && (double)score2 < 1.25 * thres) {
    return (int)((double)(float)(dist1 + dist2) / 2.0);
  }
  return 0;
}

/* gradFilter for every (dist1, dist2) in [0,255]^2 at fixed scores: out[dist1*256 + dist2]. */
void d2pc_oracle_grad_filter_table(int score1, int score2, uint8_t *out) {
  for (int a = 0; a < 256; ++a)
    for (int b = 0; b < 256; ++b)
      out[a * 256 + b] = (uint8_t)d2pc_oracle_grad_filter(a, b, score1, score2, score1, score2);
}

/* depth_map_fusion.cpp:169-217 (weightedAverage :162 truncates its weights to
 * int and divides by zero whenever both scores are >= 1; it is left out, see
 * SURVEY.md section 2). */
int d2pc_oracle_fuse_rule(int mode, int dist1, int dist2, int score1,
                          int score2) {
  switch (mode) {
    case 0:
      return d2pc_oracle_grad_filter(dist1, dist2, score1, score2, score1,
                                     score2);
    case 1: /* maxDist :169 */
      return mini(dist1, dist2);
    case 2: /* maxDistUnlessBlack :174 */
      if (dist1 == 0 || dist2 == 0) return maxi(dist1, dist2);
      return mini(dist1, dist2);
    case 3: /* betterScore :182 */
      return score1 < score2 ? dist1 : dist2;
    case 4: /* onlyGood1 :190 */
      return score2 < 50 ? dist2 : 0;
    case 5: /* onlyGoodAvg :198 */
      return (score1 < 100 && score2 < 100) ? (dist1 + dist2) / 2 : 0;
    case 6: /* overlap :205 */
      if (score1 < score2 && score1 < 20) return 150;
      if (score2 < score1 && score2 < 20) return 255;
      return 0;
    case 7: /* blackToWhite :215 */
      return 255 - score1;
    default:
      return 0;
  }
}

int d2pc_oracle_fuse(const uint8_t *d1, const uint8_t *d2, const uint8_t *s1,
                     const uint8_t *s2, int w, int h, size_t step,
                     int offset_x, int offset_y, int mode, uint8_t *fused,
                     uint8_t *combined, int dims[3]) {
  int r1[4], r2[4], rc[4];
  /* :48, :66  map/score 1: cropToSquare(image, ox, oy) */
  if (d2pc_oracle_crop_to_square(w, h, offset_x, offset_y, offset_y, r1))
    return -1;
  /* :56-57, :84-85  map/score 2: rotate (h cols x w rows), crop (-ox, -oy) */
  if (d2pc_oracle_crop_to_square(h, w, -offset_x, -offset_y, offset_y, r2))
    return -1;
  /* :105-106 container: the un-rotated frame of message 2, crop (0, 0) */
  if (d2pc_oracle_crop_to_square(w, h, 0, 0, offset_y, rc)) return -1;
  const int n = r1[2], nc = rc[2];
  if (n != r2[2] || n > nc) return -1;
  const int out_w = nc - 0 - 40, out_h = nc - 30 - 10; /* :130 cropMat */
  dims[0] = n;
  dims[1] = out_w;
  dims[2] = out_h;
  if (out_w <= 0 || out_h <= 0) return -1;

  uint8_t *rot_d2 = (uint8_t *)malloc((size_t)w * h);
  uint8_t *rot_s2 = (uint8_t *)malloc((size_t)w * h);
  d2pc_oracle_rotate_cw(d2, w, h, step, rot_d2);
  d2pc_oracle_rotate_cw(s2, w, h, step, rot_s2);

  /* container = deep copy of message 2's ROI (toCvCopy at :105) */
  uint8_t *cont = (uint8_t *)malloc((size_t)nc * nc);
  for (int i = 0; i < nc; ++i)
    memcpy(cont + (size_t)i * nc, d2 + (size_t)(rc[1] + i) * step + rc[0],
           (size_t)nc);

  /* :113-123 merge loop; score1 == grad1 == combined share one buffer (:77,
   * :113), score2 == grad2 (:96): read first, then write combined. */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      const int a1 = d1[(size_t)(r1[1] + i) * step + r1[0] + j];
      const int c1 = s1[(size_t)(r1[1] + i) * step + r1[0] + j];
      const int a2 = rot_d2[(size_t)(r2[1] + i) * h + r2[0] + j];
      const int c2 = rot_s2[(size_t)(r2[1] + i) * h + r2[0] + j];
      const int f = mode == 0
                        ? d2pc_oracle_grad_filter(a1, a2, c1, c2, c1, c2)
                        : d2pc_oracle_fuse_rule(mode, a1, a2, c1, c2);
      cont[(size_t)i * nc + j] = (uint8_t)f;
      combined[(size_t)i * n + j] = (uint8_t)mini(c1, c2);
    }
  /* :124 in-place 3x3 median on the container ROI (ROI edge = border) */
  uint8_t *med = (uint8_t *)malloc((size_t)nc * nc);
  d2pc_oracle_median_blur_u8(cont, nc, nc, (size_t)nc, med, (size_t)nc, 3);
  /* :130 cropMat(img, left 0, right 40, top 30, bottom 10) */
  for (int i = 0; i < out_h; ++i)
    memcpy(fused + (size_t)i * out_w, med + (size_t)(30 + i) * nc + 0,
           (size_t)out_w);
  free(med);
  free(cont);
  free(rot_s2);
  free(rot_d2);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* debug colouriser (depth_map_fusion.cpp:304-358)                            */
/* ------------------------------------------------------------------------- */

/* colorizeDepth: gray -> HSV rainbow (blue..red), black stays black.  The three bytes are stored in the order the
 * reference writes them (Point3_<uchar>(b, g, r)) into an image it then labels "rgb8" (:298-299). */
void d2pc_oracle_colorize_depth(const uint8_t *gray, int w, int h, size_t step, uint8_t *rgb) {
  const double maxDisp = 255;
  const float S = 1.f, V = 1.f;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const unsigned char d = (unsigned char)(40 + 0.8 * gray[(size_t)y * step + x]);
      const unsigned int H = 255 - ((unsigned char)maxDisp - d) * 280 / (unsigned char)maxDisp;
      const unsigned int hi = (H / 60) % 6;
      const float f = H / 60.f - H / 60;
      const float p = V * (1 - S), q = V * (1 - f * S), t = V * (1 - (1 - f) * S);
      float rx = 0, ry = 0, rz = 0;
      if (hi == 0) rx = p, ry = t, rz = V;
      if (hi == 1) rx = p, ry = V, rz = q;
      if (hi == 2) rx = t, ry = V, rz = p;
      if (hi == 3) rx = V, ry = q, rz = p;
      if (hi == 4) rx = V, ry = p, rz = t;
      if (hi == 5) rx = q, ry = p, rz = V;
      uint8_t *o = rgb + ((size_t)y * w + x) * 3;
      o[0] = (unsigned char)(fmaxf(0.f, fminf(rx, 1.f)) * 255.f);
      o[1] = (unsigned char)(fmaxf(0.f, fminf(ry, 1.f)) * 255.f);
      o[2] = (unsigned char)(fmaxf(0.f, fminf(rz, 1.f)) * 255.f);
      if (d == 40) o[0] = o[1] = o[2] = 0;
    }
}

/* ------------------------------------------------------------------------- */
/* matching-score preprocessing (depth_map_fusion.cpp:64-99)                  */
/* ------------------------------------------------------------------------- */

static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

/* cv::GaussianBlur on CV_8U takes OpenCV's bit-exact fixed-point path: the kernel is quantised to 8 fractional
 * bits with the rounding error diffused so it sums to 256; rows are filtered into 8.8, columns into 16.16, then
 * rounded half up.  The two kernels the reference uses (13/sigma 3, 21/sigma 10), as OpenCV 4.13 quantises them: */
static const int kGauss13[13] = {5, 8, 15, 21, 28, 33, 36, 33, 28, 21, 15, 8, 5};
static const int kGauss21[21] = {9, 9, 11, 11, 12, 13, 13, 14, 14, 15, 14, 15, 14, 14, 13, 13, 12, 11, 11, 9, 9};

/* Blurs the rectangle (x0, y0, ow, oh) of src (sw x sh); border reflect-101 at the edge of src. */
static void gauss_fixed(const uint8_t *src, int sw, int sh, size_t step, const int *k, int ks, int x0, int y0, int ow,
                        int oh, uint8_t *dst) {
  const int r = ks / 2;
  uint32_t *rows = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(oh + 2 * r) * ow);
  for (int yy = 0; yy < oh + 2 * r; ++yy) {
    const uint8_t *s = src + (size_t)reflect101(y0 + yy - r, sh) * step;
    for (int x = 0; x < ow; ++x) {
      uint32_t acc = 0;
      for (int i = 0; i < ks; ++i) acc += (uint32_t)k[i] * s[reflect101(x0 + x + i - r, sw)];
      rows[(size_t)yy * ow + x] = acc; /* 8.8 */
    }
  }
  for (int y = 0; y < oh; ++y)
    for (int x = 0; x < ow; ++x) {
      uint32_t acc = 0;
      for (int i = 0; i < ks; ++i) acc += (uint32_t)k[i] * rows[(size_t)(y + i) * ow + x];
      const uint32_t v = (acc + 32768u) >> 16;
      dst[(size_t)y * ow + x] = (uint8_t)(v > 255 ? 255 : v);
    }
  free(rows);
}

/* cv::GaussianBlur on a CV_8U SUBMATRIX with the default (non-isolated) border does not take the fixed-point path:
 * OpenCV 4.x guards it with `(borderType & BORDER_ISOLATED) || !src.isSubmatrix()` and otherwise falls through to
 * sepFilter2D(src, dst, CV_8U, kx, ky) with the float32 kernels of getGaussianKernel.  That is the call the reference
 * makes for its first blur (depth_map_fusion.cpp:70-71 / :89-90: cropped_score_k_ is mat(region), :264).  Pinned
 * against cv2.sepFilter2D 4.13.0 (AVX2 dispatch): float32 row pass (taps in order 0..12), float32 symmetric column
 * pass, then saturate_cast<uchar> (round half even); the vector loops use FMA, the scalar tails mul + add -- row
 * pass: columns >= ow - ow % 32, column pass: columns >= ow - ow % 4 (ow = ROI width).  Border: reflect-101 at
 * the edge of the PARENT image (the ROI sees the pixels around it).  getGaussianKernel(13, 3.0, CV_32F): */
static const float kGauss13f[13] = {0x1.2fd344p-6f, 0x1.17e546p-5f, 0x1.cd7846p-5f, 0x1.54699ep-4f, 0x1.c16904p-4f,
                                    0x1.09752ep-3f, 0x1.189f6cp-3f, 0x1.09752ep-3f, 0x1.c16904p-4f, 0x1.54699ep-4f,
                                    0x1.cd7846p-5f, 0x1.17e546p-5f, 0x1.2fd344p-6f};

static void gauss13_sepfilter(const uint8_t *src, int sw, int sh, size_t step, int x0, int y0, int ow, int oh,
                              uint8_t *dst, size_t dst_step) {
  const int r = 6, row_vec_end = ow - ow % 32, col_vec_end = ow - ow % 4;
  const float *k = kGauss13f;
  float *t = (float *)malloc(sizeof(float) * (size_t)(oh + 2 * r) * ow);
  for (int yy = 0; yy < oh + 2 * r; ++yy) {
    const uint8_t *s = src + (size_t)reflect101(y0 + yy - r, sh) * step;
    for (int x = 0; x < ow; ++x) {
      float acc = k[0] * (float)s[reflect101(x0 + x - r, sw)];
      for (int i = 1; i < 13; ++i) {
        const float v = (float)s[reflect101(x0 + x + i - r, sw)];
        acc = x < row_vec_end ? fmaf(k[i], v, acc) : acc + k[i] * v;
      }
      t[(size_t)yy * ow + x] = acc;
    }
  }
  for (int y = 0; y < oh; ++y)
    for (int x = 0; x < ow; ++x) {
      float acc = k[r] * t[(size_t)(y + r) * ow + x];
      for (int j = 1; j <= r; ++j) {
        const float v = t[(size_t)(y + r + j) * ow + x] + t[(size_t)(y + r - j) * ow + x];
        acc = x < col_vec_end ? fmaf(k[r + j], v, acc) : acc + k[r + j] * v;
      }
      const float q = rintf(acc);
      dst[(size_t)y * dst_step + x] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
    }
  free(t);
}

/* cv::GaussianBlur(roi, dst, Size(ksize, ksize), sigma) for the two calls the reference makes (13 / 3.0, 21 / 10.0),
 * roi = rect of a parent image; `submatrix` is Mat::isSubmatrix() of the source (rect smaller than the parent). */
int d2pc_oracle_gaussian_blur_u8(const uint8_t *parent, int pw, int ph, size_t step, const int rect[4], int ksize,
                                 double sigma, int submatrix, uint8_t *dst, size_t dst_step) {
  const int x0 = rect[0], y0 = rect[1], ow = rect[2], oh = rect[3];
  if (ow <= 0 || oh <= 0 || x0 < 0 || y0 < 0 || x0 + ow > pw || y0 + oh > ph) return -1;
  const int is13 = ksize == 13 && sigma == 3.0, is21 = ksize == 21 && sigma == 10.0;
  if (!is13 && !is21) return -2;
  if (submatrix) {
    if (!is13) return -2;
    gauss13_sepfilter(parent, pw, ph, step, x0, y0, ow, oh, dst, dst_step);
    return 0;
  }
  /* not a submatrix: the fixed-point path; rect is the whole image */
  uint8_t *tmp = (uint8_t *)malloc((size_t)ow * oh);
  gauss_fixed(parent, pw, ph, step, is13 ? kGauss13 : kGauss21, ksize, x0, y0, ow, oh, tmp);
  for (int y = 0; y < oh; ++y) memcpy(dst + (size_t)y * dst_step, tmp + (size_t)y * ow, (size_t)ow);
  free(tmp);
  return 0;
}

/* cv::Sobel(src, dst, -1, dx, dy, 7, 0.03) on an n x n CV_8U image, (dx,dy) = (0,2) or (2,0): separable float32
 * filter, kernels getDerivKernels(7): smoothing {1,6,15,20,15,6,1}, 2nd derivative {1,2,-1,-4,-1,2,1}; the
 * smoothing kernel is the one scaled (in float32).  The pass with the scaled kernel rounds at every step; OpenCV's
 * AVX2 code uses FMA in its vector loop and mul+add in the scalar tail: the row pass vectorises 32 columns, the
 * symmetric column pass 4 columns.  The other pass works on integers < 2^24 and is exact. */
static void sobel7_second(const uint8_t *src, int n, int vertical, uint8_t *dst) {
  static const float smooth[7] = {1, 6, 15, 20, 15, 6, 1}, deriv[7] = {1, 2, -1, -4, -1, 2, 1};
  float ks[7];
  for (int i = 0; i < 7; ++i) ks[i] = smooth[i] * 0.03f;
  float *t = (float *)malloc(sizeof(float) * (size_t)n * n);
  if (!vertical) {
    /* dx = 0, dy = 2: rows = scaled smoothing (sequential k = 0..6), columns = integer derivative */
    const int vec_end = n - n % 32;
    for (int y = 0; y < n; ++y)
      for (int x = 0; x < n; ++x) {
        float s = ks[0] * (float)src[(size_t)y * n + reflect101(x - 3, n)];
        for (int i = 1; i < 7; ++i) {
          const float v = (float)src[(size_t)y * n + reflect101(x + i - 3, n)];
          s = x < vec_end ? fmaf(ks[i], v, s) : s + ks[i] * v;
        }
        t[(size_t)y * n + x] = s;
      }
    for (int y = 0; y < n; ++y)
      for (int x = 0; x < n; ++x) {
        float s = deriv[3] * t[(size_t)y * n + x];
        for (int i = 1; i <= 3; ++i)
          s += deriv[3 + i] * (t[(size_t)reflect101(y + i, n) * n + x] + t[(size_t)reflect101(y - i, n) * n + x]);
        const float r = rintf(s);
        dst[(size_t)y * n + x] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
      }
  } else {
    /* dx = 2, dy = 0: rows = integer derivative (exact), columns = scaled smoothing, symmetric form */
    const int vec_end = n - n % 4;
    for (int y = 0; y < n; ++y)
      for (int x = 0; x < n; ++x) {
        float s = 0.f;
        for (int i = 0; i < 7; ++i) s += deriv[i] * (float)src[(size_t)y * n + reflect101(x + i - 3, n)];
        t[(size_t)y * n + x] = s;
      }
    for (int y = 0; y < n; ++y)
      for (int x = 0; x < n; ++x) {
        float s = ks[3] * t[(size_t)y * n + x];
        for (int i = 1; i <= 3; ++i) {
          const float v = t[(size_t)reflect101(y + i, n) * n + x] + t[(size_t)reflect101(y - i, n) * n + x];
          s = x < vec_end ? fmaf(ks[3 + i], v, s) : s + ks[3 + i] * v;
        }
        const float r = rintf(s);
        dst[(size_t)y * n + x] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
      }
  }
  free(t);
}

/* cv::Sobel(src, dst, -1, dx, dy, 7, 0.03) for (dx, dy) = (0, 2) [vertical = 0] or (2, 0) [vertical = 1] on a dense
 * n x n CV_8U image (exported for the reference harness, oracle/ref_stubs). */
void d2pc_oracle_sobel7_second_u8(const uint8_t *src, int n, int vertical, uint8_t *dst) {
  sobel7_second(src, n, vertical, dst);
}

int d2pc_oracle_score_preprocess(const uint8_t *frame, int w, int h, size_t step, const int rect[4], int vertical,
                                 uint8_t *out) {
  const int x0 = rect[0], y0 = rect[1], n = rect[2];
  if (n <= 0 || rect[3] != n || x0 < 0 || y0 < 0 || x0 + n > w || y0 + n > h) return -1;
  uint8_t *g = (uint8_t *)malloc((size_t)n * n), *e = (uint8_t *)malloc((size_t)n * n);
  /* :70-71 / :89-90: the source is the ROI mat(region) (:264), a submatrix unless the region is the whole frame */
  d2pc_oracle_gaussian_blur_u8(frame, w, h, step, rect, 13, 3.0, n < w || n < h, g, (size_t)n);
  sobel7_second(g, n, vertical, e);                                            /* :72 / :91 */
  for (size_t i = 0; i < (size_t)n * n; ++i) e[i] = e[i] > 30 ? 255 : 0;       /* :73 / :92 THRESH_BINARY */
  gauss_fixed(e, n, n, (size_t)n, kGauss21, 21, 0, 0, n, n, g);              /* :74-75 / :93-94 */
  for (int y = 0; y < n; ++y)                                                  /* :76 / :95 saturating s + 2g */
    for (int x = 0; x < n; ++x) {
      const int v = frame[(size_t)(y0 + y) * step + x0 + x] + 2 * g[(size_t)y * n + x];
      out[(size_t)y * n + x] = (uint8_t)(v > 255 ? 255 : v);
    }
  free(e);
  free(g);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* CPU baseline driver                                                        */
/* ------------------------------------------------------------------------- */

typedef struct {
  const uint8_t *frames;
  int n_frames, w, h, mono8, tid, n_threads;
  size_t step, cloud_stride;
  const double *q;
  uint8_t *cloud;
  size_t points;
} run_arg;

static void *run_worker(void *p) {
  run_arg *a = (run_arg *)p;
  const size_t fstride = a->step * (size_t)a->h;
  for (int i = a->tid; i < a->n_frames; i += a->n_threads) {
    const uint8_t *f = a->frames + fstride * i;
    uint8_t *c = a->cloud + a->cloud_stride * i;
    a->points += a->mono8
                     ? d2pc_oracle_disparity_cb_mono8(f, a->w, a->h, a->step,
                                                      a->q, c)
                     : d2pc_oracle_disparity_cb_f32((const float *)f, a->w,
                                                    a->h, a->step, a->q, c);
  }
  return NULL;
}

size_t d2pc_oracle_run_frames(const void *frames, int n_frames, int w, int h,
                              size_t step, int mono8, const double q[16],
                              uint8_t *cloud, size_t cloud_stride,
                              int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  run_arg args[256];
  pthread_t th[256];
  for (int t = 0; t < n_threads; ++t) {
    run_arg a = {(const uint8_t *)frames, n_frames, w, h, mono8, t, n_threads,
                 step, cloud_stride, q, cloud, 0};
    args[t] = a;
  }
  if (n_threads == 1) {
    run_worker(&args[0]);
    return args[0].points;
  }
  for (int t = 0; t < n_threads; ++t)
    pthread_create(&th[t], NULL, run_worker, &args[t]);
  size_t total = 0;
  for (int t = 0; t < n_threads; ++t) {
    pthread_join(th[t], NULL);
    total += args[t].points;
  }
  return total;
}
