// ref_harness.cpp -- TEST INFRASTRUCTURE: C entry points over the reference's own classes, compiled together with
// /root/reference/src/depth_map_fusion.cpp and /root/reference/src/disparity_to_point_cloud.cpp (unmodified, from
// where they lie) against oracle/ref_stubs/ into oracle/_ref/libd2pc_ref.so.  See oracle/ref_stubs/ref_stubs.hpp
// for what is reference code and what is a stand-in.  Only tests/ load this library.
#include <cstring>
#include <memory>

#include "disparity_to_point_cloud/depth_map_fusion.hpp"
#include "disparity_to_point_cloud/disparity_to_point_cloud.hpp"

using depth_map_fusion::DepthMapFusion;
using d2pc::Disparity2PCloud;

namespace {
sensor_msgs::ImageConstPtr make_image(const uint8_t *img, int w, int h, int step, const char *encoding, uint32_t seq,
                                      uint32_t sec, uint32_t nsec, const char *frame_id) {
  auto m = std::make_shared<sensor_msgs::Image>();
  m->header.seq = seq, m->header.stamp.sec = sec, m->header.stamp.nsec = nsec;
  m->header.frame_id = frame_id ? frame_id : "";
  m->height = (uint32_t)h, m->width = (uint32_t)w, m->step = (uint32_t)step;
  m->encoding = encoding ? encoding : "mono8";
  m->data.assign(img, img + (size_t)step * h);
  return m;
}
template <class F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const cv_bridge::Exception &) {
    return -2;  // the reference would terminate on the uncaught exception
  } catch (const cv::Exception &) {
    return -3;
  }
}
}  // namespace

extern "C" {

// ---- captured ROS traffic ---------------------------------------------------------------------------------
void ref_reset(void) { ref_capture::state() = ref_capture::State(); }
void ref_set_param(const char *name, double v) { ref_capture::state().params[name] = v; }
void ref_published_clear(void) { ref_capture::state().published.clear(); }
int ref_published_count(void) { return (int)ref_capture::state().published.size(); }
int ref_warnings(void) { return ref_capture::state().warnings; }
// info = {is_cloud, width, height, step | point_step, latched, queue, seq, sec, nsec, is_dense, row_step, n_fields}
int ref_published_info(int idx, char topic[64], char encoding_or_frame[64], uint32_t info[12]) {
  const auto &v = ref_capture::state().published;
  if (idx < 0 || idx >= (int)v.size()) return -1;
  const auto &p = v[idx];
  strncpy(topic, p.topic.c_str(), 63), topic[63] = 0;
  memset(info, 0, sizeof(uint32_t) * 12);
  info[0] = p.is_cloud, info[4] = p.latched, info[5] = p.queue;
  if (p.is_cloud) {
    const auto &c = p.cloud;
    strncpy(encoding_or_frame, c.header.frame_id.c_str(), 63);
    info[1] = c.width, info[2] = c.height, info[3] = c.point_step, info[6] = c.header.seq, info[7] = c.header.stamp.sec,
    info[8] = c.header.stamp.nsec, info[9] = c.is_dense, info[10] = c.row_step, info[11] = (uint32_t)c.fields.size();
  } else {
    const auto &m = p.image;
    strncpy(encoding_or_frame, m.encoding.c_str(), 63);
    info[1] = m.width, info[2] = m.height, info[3] = m.step, info[6] = m.header.seq, info[7] = m.header.stamp.sec,
    info[8] = m.header.stamp.nsec;
  }
  encoding_or_frame[63] = 0;
  return 0;
}
size_t ref_published_data(int idx, uint8_t *out, size_t cap) {
  const auto &v = ref_capture::state().published;
  if (idx < 0 || idx >= (int)v.size()) return 0;
  const std::vector<uint8_t> &d = v[idx].is_cloud ? v[idx].cloud.data : v[idx].image.data;
  if (out && cap >= d.size() && !d.empty()) memcpy(out, d.data(), d.size());
  return d.size();
}
// field k of a published cloud: name, {offset, datatype, count}
int ref_published_field(int idx, int k, char name[16], uint32_t odc[3]) {
  const auto &v = ref_capture::state().published;
  if (idx < 0 || idx >= (int)v.size() || !v[idx].is_cloud || k < 0 || k >= (int)v[idx].cloud.fields.size()) return -1;
  const auto &f = v[idx].cloud.fields[k];
  strncpy(name, f.name.c_str(), 15), name[15] = 0;
  odc[0] = f.offset, odc[1] = f.datatype, odc[2] = f.count;
  return 0;
}
int ref_topics(int advertised, int idx, char topic[64], uint32_t queue_latch[2]) {
  auto &s = ref_capture::state();
  const int n = advertised ? (int)s.advertised.size() : (int)s.subscribed.size();
  if (idx < 0) return n;
  if (idx >= n) return -1;
  if (advertised) {
    strncpy(topic, s.advertised[idx].topic.c_str(), 63);
    queue_latch[0] = s.advertised[idx].queue, queue_latch[1] = s.advertised[idx].latch;
  } else {
    strncpy(topic, s.subscribed[idx].topic.c_str(), 63);
    queue_latch[0] = s.subscribed[idx].queue, queue_latch[1] = 0;
  }
  topic[63] = 0;
  return n;
}

// ---- DepthMapFusion (include/disparity_to_point_cloud/depth_map_fusion.hpp:63-155) -----------------------------
void *ref_fusion_create(void) { return new DepthMapFusion(); }  // reads ~offset_x / ~offset_y (ref_set_param)
void ref_fusion_destroy(void *n) { delete static_cast<DepthMapFusion *>(n); }
void ref_fusion_offsets(void *n, int out[2]) {
  out[0] = static_cast<DepthMapFusion *>(n)->offset_x_, out[1] = static_cast<DepthMapFusion *>(n)->offset_y_;
}
// which: 1 DisparityCb1, 2 DisparityCb2, 3 MatchingScoreCb1, 4 MatchingScoreCb2 (src/depth_map_fusion.cpp:46-99)
int ref_fusion_callback(void *node, int which, const uint8_t *img, int w, int h, int step, const char *encoding,
                        uint32_t seq, uint32_t sec, uint32_t nsec) {
  DepthMapFusion *n = static_cast<DepthMapFusion *>(node);
  const sensor_msgs::ImageConstPtr m = make_image(img, w, h, step, encoding, seq, sec, nsec, "cam");
  return guarded([&] {
    if (which == 1) n->DisparityCb1(m);
    else if (which == 2) n->DisparityCb2(m);
    else if (which == 3) n->MatchingScoreCb1(m);
    else n->MatchingScoreCb2(m);
  });
}
int ref_grad_filter(void *n, int d1, int d2, int s1, int s2, int g1, int g2) {
  return static_cast<DepthMapFusion *>(n)->gradFilter(d1, d2, s1, s2, g1, g2);
}
// mode numbering of include/d2pc_b200.h d2pc_fuse_rule (1..7); 8 = weightedAverage (divides by zero for scores >= 2)
int ref_fuse_rule(void *node, int mode, int d1, int d2, int s1, int s2) {
  DepthMapFusion *n = static_cast<DepthMapFusion *>(node);
  switch (mode) {
    case 1: return n->maxDist(d1, d2, s1, s2);
    case 2: return n->maxDistUnlessBlack(d1, d2, s1, s2);
    case 3: return n->betterScore(d1, d2, s1, s2);
    case 4: return n->onlyGood1(d1, d2, s1, s2);
    case 5: return n->onlyGoodAvg(d1, d2, s1, s2);
    case 6: return n->overlap(d1, d2, s1, s2);
    case 7: return n->blackToWhite(d1, d2, s1, s2);
    case 8: return n->weightedAverage(d1, d2, s1, s2);
    default: return -1;
  }
}
// cropToSquare / cropMat on a cols x rows image: the rectangle of the returned ROI {x, y, w, h}; -3 if cv::Mat throws
int ref_crop_to_square(void *node, int cols, int rows, int offset_x, int offset_y, int rect[4]) {
  return guarded([&] {
    cv::Mat m(rows, cols, CV_8UC1);
    const cv::Mat r = static_cast<DepthMapFusion *>(node)->cropToSquare(m, offset_x, offset_y);
    rect[0] = r.ofs_x, rect[1] = r.ofs_y, rect[2] = r.cols, rect[3] = r.rows;
  });
}
int ref_crop_mat(void *node, int cols, int rows, int left, int right, int top, int bottom, int rect[4]) {
  return guarded([&] {
    cv::Mat m(rows, cols, CV_8UC1);
    const cv::Mat r = static_cast<DepthMapFusion *>(node)->cropMat(m, left, right, top, bottom);
    rect[0] = r.ofs_x, rect[1] = r.ofs_y, rect[2] = r.cols, rect[3] = r.rows;
  });
}
void ref_rotate(void *node, const uint8_t *img, int w, int h, uint8_t *out /* h cols x w rows */) {
  cv::Mat m(h, w, CV_8UC1);
  memcpy(m.data, img, (size_t)w * h);
  const cv::Mat r = static_cast<DepthMapFusion *>(node)->rotateMat(m);
  for (int i = 0; i < r.rows; ++i) memcpy(out + (size_t)i * r.cols, r.data + (size_t)i * r.step, (size_t)r.cols);
}
void ref_colorize(void *node, const uint8_t *gray, int w, int h, uint8_t *rgb) {
  cv::Mat g(h, w, CV_8UC1), c;
  memcpy(g.data, gray, (size_t)w * h);
  static_cast<DepthMapFusion *>(node)->colorizeDepth(g, c);
  for (int i = 0; i < h; ++i) memcpy(rgb + (size_t)i * w * 3, c.data + (size_t)i * c.step, (size_t)w * 3);
}

// ---- Disparity2PCloud (include/disparity_to_point_cloud/disparity_to_point_cloud.hpp:60-109) -----------------------
void *ref_d2pc_create(void) { return new Disparity2PCloud(); }  // reads ~fx_ ~fy_ ~cx_ ~cy_ ~base_line_
void ref_d2pc_destroy(void *n) { delete static_cast<Disparity2PCloud *>(n); }
int ref_d2pc_callback(void *node, const uint8_t *img, int w, int h, int step, const char *encoding, uint32_t seq,
                      uint32_t sec, uint32_t nsec) {
  const sensor_msgs::ImageConstPtr m = make_image(img, w, h, step, encoding, seq, sec, nsec, "cam");
  return guarded([&] { static_cast<Disparity2PCloud *>(node)->DisparityCb(m); });
}

}  // extern "C"
