// ref_stubs.hpp -- TEST INFRASTRUCTURE (part of oracle/, never linked into the product).
//
// Minimal stand-ins for the third-party headers the reference includes (ros/ros.h, cv_bridge, opencv2/*, pcl/*,
// sensor_msgs/*), just enough for the reference's own translation units
//     /root/reference/src/depth_map_fusion.cpp
//     /root/reference/src/disparity_to_point_cloud.cpp
// to compile UNMODIFIED, from where they lie, into oracle/_ref/libd2pc_ref.so (recipe: oracle/Makefile).  What is
// compiled from the reference is all of its own logic: the callback sequencing and the cv::Mat aliasing between
// callbacks, cropToSquare / cropMat / rotateMat, gradFilter and the seven alternate rules, the merge loop,
// colorizeDepth, the crop / push_back loop and the cloud metadata of DisparityCb.  What the stubs supply:
//   * cv::Mat with OpenCV's sharing semantics (reference-counted buffer, ROI views that alias their parent,
//     create() that keeps a matching allocation, the SUBMATRIX flag rule of Mat(const Mat&, const Rect&));
//   * the OpenCV *functions* (medianBlur, GaussianBlur, Sobel, threshold, reprojectImageTo3D, stereoRectify,
//     convertTo, transpose, flip, saturating MatExpr a + s*b), bound to the cv2-pinned C restatement
//     oracle/d2pc_oracle.c -- third-party arithmetic is pinned against cv2, not here;
//   * ROS plumbing that records what a node publishes instead of sending it;
//   * pcl::PointXYZ / PointCloud / toROSMsg restated from PCL's public definitions.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../d2pc_oracle.h"

namespace boost {
template <class T>
using shared_ptr = std::shared_ptr<T>;
}

// ---------------------------------------------------------------------------------------------------------
// messages
// ---------------------------------------------------------------------------------------------------------
namespace ros {
struct Time {
  uint32_t sec = 0, nsec = 0;
};
}  // namespace ros
namespace std_msgs {
struct Header {
  uint32_t seq = 0;
  ros::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs
namespace sensor_msgs {
struct Image {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::string encoding;
  uint8_t is_bigendian = 0;
  uint32_t step = 0;
  std::vector<uint8_t> data;
};
typedef boost::shared_ptr<Image> ImagePtr;
typedef boost::shared_ptr<Image const> ImageConstPtr;
struct PointField {
  enum { INT8 = 1, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 };
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 0;
};
struct PointCloud2 {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  uint8_t is_bigendian = 0;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  uint8_t is_dense = 0;
};
}  // namespace sensor_msgs

// ---------------------------------------------------------------------------------------------------------
// ROS plumbing: publishers record, parameters come from a table the harness fills
// ---------------------------------------------------------------------------------------------------------
namespace ref_capture {
struct Published {
  std::string topic;
  bool latched = false;
  uint32_t queue = 0;
  bool is_cloud = false;
  sensor_msgs::Image image;
  sensor_msgs::PointCloud2 cloud;
};
struct Advertised {
  std::string topic;
  uint32_t queue;
  bool latch;
};
struct Subscribed {
  std::string topic;
  uint32_t queue;
};
struct State {
  std::vector<Published> published;
  std::vector<Advertised> advertised;
  std::vector<Subscribed> subscribed;
  std::map<std::string, double> params;
  int warnings = 0;
};
inline State &state() {
  static State s;
  return s;
}
}  // namespace ref_capture

#define ROS_WARN(...) (++ref_capture::state().warnings)

namespace ros {
class Subscriber {};
class Publisher {
 public:
  std::string topic;
  uint32_t queue = 0;
  bool latch = false;
  void publish(const sensor_msgs::Image &m) const {
    ref_capture::Published p;
    p.topic = topic, p.latched = latch, p.queue = queue, p.image = m;
    ref_capture::state().published.push_back(p);
  }
  void publish(const sensor_msgs::ImagePtr &m) const { publish(*m); }
  void publish(const sensor_msgs::PointCloud2 &m) const {
    ref_capture::Published p;
    p.topic = topic, p.latched = latch, p.queue = queue, p.is_cloud = true, p.cloud = m;
    ref_capture::state().published.push_back(p);
  }
};
class NodeHandle {
 public:
  explicit NodeHandle(const std::string & = std::string()) {}
  template <class M, class T>
  Subscriber subscribe(const std::string &topic, uint32_t queue, void (T::*)(const boost::shared_ptr<M const> &), T *) {
    ref_capture::state().subscribed.push_back({topic, queue});
    return Subscriber();
  }
  template <class M>
  Publisher advertise(const std::string &topic, uint32_t queue, bool latch = false) {
    Publisher p;
    p.topic = topic, p.queue = queue, p.latch = latch;
    ref_capture::state().advertised.push_back({topic, queue, latch});
    return p;
  }
  bool getParam(const std::string &name, int &v) const {
    auto it = ref_capture::state().params.find(name);
    if (it == ref_capture::state().params.end()) return false;
    v = (int)it->second;
    return true;
  }
  template <class T>
  void param(const std::string &name, T &v, const T &def) const {
    auto it = ref_capture::state().params.find(name);
    v = it == ref_capture::state().params.end() ? def : (T)it->second;
  }
};
}  // namespace ros

// ---------------------------------------------------------------------------------------------------------
// cv:: -- Mat with OpenCV's sharing semantics; functions bound to the cv2-pinned C restatement
// ---------------------------------------------------------------------------------------------------------
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {
typedef unsigned char uchar;
struct Exception : std::runtime_error {
  explicit Exception(const std::string &w) : std::runtime_error(w) {}
};
struct Size {
  int width = 0, height = 0;
  Size() {}
  Size(int w, int h) : width(w), height(h) {}
  bool operator==(const Size &o) const { return width == o.width && height == o.height; }
};
struct Rect {
  int x = 0, y = 0, width = 0, height = 0;
  Rect() {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
template <class T>
struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
};
typedef Point3_<float> Point3f;
template <class T, int N>
struct Vec {
  T val[N];
  T &operator[](int i) { return val[i]; }
  const T &operator[](int i) const { return val[i]; }
};
typedef Vec<float, 3> Vec3f;
struct Scalar {
  double v[4];
  static Scalar all(double a) {
    Scalar s;
    s.v[0] = s.v[1] = s.v[2] = s.v[3] = a;
    return s;
  }
};

class Mat;
struct MatExpr {  // a + alpha * b (either may be absent), the only expression the reference builds
  const Mat *a = nullptr, *b = nullptr;
  double alpha = 1.0;
};

class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uchar *data = nullptr;
  // the allocation this header views (shared, like OpenCV's refcount) and its full extent
  std::shared_ptr<std::vector<uchar>> buf;
  int whole_rows = 0, whole_cols = 0, ofs_x = 0, ofs_y = 0;
  bool submatrix = false;

  Mat() {}
  Mat(Size s, int type) { create(s.height, s.width, type); }
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(const Mat &m, const Rect &roi) : rows(roi.height), cols(roi.width), step(m.step), buf(m.buf), type_(m.type_) {
    // modules/core/src/matrix.cpp Mat::Mat(const Mat&, const Rect&): bounds assert, then the SUBMATRIX flag rule
    if (!(0 <= roi.x && 0 <= roi.width && roi.x + roi.width <= m.cols && 0 <= roi.y && 0 <= roi.height &&
          roi.y + roi.height <= m.rows))
      throw Exception("Mat(roi): rectangle outside the matrix");
    data = m.data + (size_t)roi.y * m.step + (size_t)roi.x * m.elemSize();
    whole_rows = m.whole_rows, whole_cols = m.whole_cols, ofs_x = m.ofs_x + roi.x, ofs_y = m.ofs_y + roi.y;
    submatrix = m.submatrix || roi.width < m.cols || roi.height < m.rows;
  }
  Mat operator()(const Rect &roi) const { return Mat(*this, roi); }

  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize() const {
    const int d = depth();
    return (size_t)channels() * (d == CV_8U ? 1 : d == CV_32F ? 4 : 8);
  }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  Size size() const { return Size(cols, rows); }
  bool isSubmatrix() const { return submatrix; }
  bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }

  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;  // Mat::create keeps a matching allocation
    type_ = type;
    rows = r, cols = c;
    step = (size_t)c * elemSize();
    buf = std::make_shared<std::vector<uchar>>((size_t)r * step + 64);
    data = buf->data();
    whole_rows = r, whole_cols = c, ofs_x = ofs_y = 0;
    submatrix = false;
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int i = 0; i < rows; ++i) memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, (size_t)cols * elemSize());
    return m;
  }
  static Mat eye(int r, int c, int type) {
    Mat m(r, c, type);
    memset(m.data, 0, (size_t)r * m.step);
    for (int i = 0; i < (r < c ? r : c); ++i) m.at<double>(i, i) = 1.0;
    return m;
  }
  template <class T>
  T &at(int i, int j) {
    return *reinterpret_cast<T *>(data + (size_t)i * step + (size_t)j * sizeof(T));
  }
  template <class T>
  const T &at(int i, int j) const {
    return *reinterpret_cast<const T *>(data + (size_t)i * step + (size_t)j * sizeof(T));
  }
  template <class T>
  T *ptr(int i) {
    return reinterpret_cast<T *>(data + (size_t)i * step);
  }
  Mat &operator=(const Scalar &s) {
    for (int i = 0; i < rows; ++i)
      for (size_t b = 0; b < (size_t)cols * elemSize(); ++b) data[(size_t)i * step + b] = (uchar)s.v[0];
    return *this;
  }
  Mat &operator=(const MatExpr &e);  // evaluates into *this via create() (in place when size and type match)
  void convertTo(Mat &dst, int rtype, double alpha = 1.0, double beta = 0.0) const {
    if (depth() != CV_8U || rtype != CV_32FC1 || beta != 0.0) throw Exception("convertTo: only 8U -> 32F is bound");
    dst.create(rows, cols, CV_32FC1);
    d2pc_oracle_convert_u8_f32(data, cols, rows, step, reinterpret_cast<float *>(dst.data), dst.step, alpha);
  }
  // the full image a ROI lives in (Mat::locateROI): what non-isolated border handling reads
  const uchar *whole_data() const { return data - (size_t)ofs_y * step - (size_t)ofs_x * elemSize(); }

 private:
  int type_ = 0;
};

inline MatExpr operator*(double s, const Mat &b) {
  MatExpr e;
  e.b = &b, e.alpha = s;
  return e;
}
inline MatExpr operator+(const Mat &a, const MatExpr &e) {
  MatExpr r = e;
  r.a = &a;
  return r;
}
inline Mat &Mat::operator=(const MatExpr &e) {
  // cv::add / scaleAdd on CV_8U: saturate_cast<uchar>(a + alpha * b); alpha is integral here, so exact
  const Mat a = *e.a, b = *e.b;  // keep the operands alive: *this may be one of them
  if (a.type() != CV_8UC1 || b.type() != CV_8UC1 || !(a.size() == b.size())) throw Exception("MatExpr: 8UC1 only");
  create(a.rows, a.cols, CV_8UC1);
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < cols; ++j) {
      const double v = a.at<uchar>(i, j) + e.alpha * b.at<uchar>(i, j);
      const long r = lrint(v);
      at<uchar>(i, j) = (uchar)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
  return *this;
}

template <class T>
class Mat_ : public Mat {
 public:
  Mat_(int r, int c) : Mat(r, c, sizeof(T) == 8 ? CV_64FC1 : CV_32FC1) {}
};
template <class T>
struct MatCommaInitializer_ {
  Mat_<T> m;
  int idx;
  MatCommaInitializer_(const Mat_<T> &mm, T first) : m(mm), idx(0) { put(first); }
  void put(T v) {
    m.template at<T>(idx / m.cols, idx % m.cols) = v;
    ++idx;
  }
  MatCommaInitializer_ &operator,(T v) {
    put(v);
    return *this;
  }
  operator Mat() const { return m; }
};
template <class T>
MatCommaInitializer_<T> operator<<(const Mat_<T> &m, T v) {
  return MatCommaInitializer_<T>(m, v);
}
template <class T>
MatCommaInitializer_<T> operator<<(const Mat_<T> &m, int v) {
  return MatCommaInitializer_<T>(m, (T)v);
}

// ---- imgproc / core / calib3d functions the reference calls ------------------------------------------------
inline void medianBlur(const Mat &src, Mat &dst, int ksize) {
  if (src.type() != CV_8UC1) throw Exception("medianBlur: 8UC1 only");
  const Mat s = src.data == dst.data ? src.clone() : src;  // in-place call: OpenCV filters from a copy
  dst.create(s.rows, s.cols, CV_8UC1);
  d2pc_oracle_median_blur_u8(s.data, s.cols, s.rows, s.step, dst.data, dst.step, ksize);  // replicate at the ROI edge
}
inline void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigma1, double sigma2 = 0, int borderType = 4) {
  if (src.type() != CV_8UC1 || ksize.width != ksize.height || (sigma2 != 0 && sigma2 != sigma1) || borderType != 4)
    throw Exception("GaussianBlur: unsupported arguments");
  const Mat s = src;
  const bool inplace = s.data == dst.data;
  dst.create(s.rows, s.cols, CV_8UC1);
  Mat out = inplace ? Mat(s.rows, s.cols, CV_8UC1) : dst;
  const int rect[4] = {s.ofs_x, s.ofs_y, s.cols, s.rows};
  // smooth.dispatch.cpp: the fixed-point path needs BORDER_ISOLATED or !src.isSubmatrix(); otherwise sepFilter2D,
  // whose non-isolated border reads the pixels around the ROI (locateROI)
  if (d2pc_oracle_gaussian_blur_u8(s.whole_data(), s.whole_cols, s.whole_rows, s.step, rect, ksize.width, sigma1,
                                   s.isSubmatrix() ? 1 : 0, out.data, out.step))
    throw Exception("GaussianBlur: kernel not bound");
  if (inplace)
    for (int i = 0; i < s.rows; ++i) memcpy(dst.data + (size_t)i * dst.step, out.data + (size_t)i * out.step, (size_t)s.cols);
}
inline void Sobel(const Mat &src, Mat &dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0,
                  int borderType = 4) {
  const bool ok = src.type() == CV_8UC1 && ddepth == -1 && ksize == 7 && scale == 0.03 && delta == 0 && borderType == 4 &&
                  ((dx == 0 && dy == 2) || (dx == 2 && dy == 0)) && src.rows == src.cols && !src.isSubmatrix();
  if (!ok) throw Exception("Sobel: unsupported arguments");
  const Mat s = src.clone();
  dst.create(s.rows, s.cols, CV_8UC1);
  std::vector<uchar> tmp((size_t)s.rows * s.cols);
  d2pc_oracle_sobel7_second_u8(s.data, s.rows, dx == 2 ? 1 : 0, tmp.data());
  for (int i = 0; i < s.rows; ++i) memcpy(dst.data + (size_t)i * dst.step, tmp.data() + (size_t)i * s.cols, (size_t)s.cols);
}
inline double threshold(const Mat &src, Mat &dst, double thresh, double maxval, int type) {
  if (src.type() != CV_8UC1 || type != 0) throw Exception("threshold: THRESH_BINARY on 8UC1 only");
  const Mat s = src;
  dst.create(s.rows, s.cols, CV_8UC1);
  const int t = (int)std::floor(thresh);  // imgproc/thresh.cpp: for CV_8U thresh is floored, maxval rounded
  const long mv = lrint(maxval);
  for (int i = 0; i < s.rows; ++i)
    for (int j = 0; j < s.cols; ++j) dst.at<uchar>(i, j) = s.at<uchar>(i, j) > t ? (uchar)(mv > 255 ? 255 : mv) : 0;
  return thresh;
}
inline void transpose(const Mat &src, Mat &dst) {
  const Mat s = src.data == dst.data ? src.clone() : src;
  dst.create(s.cols, s.rows, s.type());
  for (int i = 0; i < s.rows; ++i)
    for (int j = 0; j < s.cols; ++j) dst.at<uchar>(j, i) = s.at<uchar>(i, j);
}
inline void flip(const Mat &src, Mat &dst, int flipCode) {
  if (flipCode != 1) throw Exception("flip: only around the y axis");
  const Mat s = src.data == dst.data ? src.clone() : src;
  dst.create(s.rows, s.cols, s.type());
  for (int i = 0; i < s.rows; ++i)
    for (int j = 0; j < s.cols; ++j) dst.at<uchar>(i, j) = s.at<uchar>(i, s.cols - 1 - j);
}
inline void applyColorMap(const Mat &, Mat &, int) { throw Exception("applyColorMap: not reached by the reference"); }
inline void reprojectImageTo3D(const Mat &disp, Mat &out, const Mat &Q, bool handleMissingValues = false, int ddepth = -1) {
  if (disp.type() != CV_32FC1 || Q.type() != CV_64FC1 || Q.rows != 4 || Q.cols != 4 || handleMissingValues || ddepth != -1)
    throw Exception("reprojectImageTo3D: unsupported arguments");
  out.create(disp.rows, disp.cols, CV_32FC3);
  double q[16];
  for (int i = 0; i < 16; ++i) q[i] = Q.at<double>(i / 4, i % 4);
  d2pc_oracle_reproject_image_to_3d(reinterpret_cast<const float *>(disp.data), disp.cols, disp.rows, disp.step, q,
                                    reinterpret_cast<float *>(out.data));
}
inline void stereoRectify(const Mat &K1, const Mat &D1, const Mat &K2, const Mat &D2, Size size, const Mat &R, const Mat &t,
                          Mat &R1, Mat &R2, Mat &P1, Mat &P2, Mat &Q) {
  // bound for the reference's call only: identical cameras, zero distortion, R = I, t = (-b, 0, 0)
  for (int i = 0; i < 5; ++i)
    if (D1.at<double>(i, 0) != 0 || D2.at<double>(i, 0) != 0) throw Exception("stereoRectify: distortion not bound");
  for (int i = 0; i < 9; ++i)
    if (K1.at<double>(i / 3, i % 3) != K2.at<double>(i / 3, i % 3) || R.at<double>(i / 3, i % 3) != (i % 4 == 0 ? 1.0 : 0.0))
      throw Exception("stereoRectify: only K1 == K2, R = I is bound");
  if (t.at<double>(1, 0) != 0 || t.at<double>(2, 0) != 0) throw Exception("stereoRectify: only a horizontal rig is bound");
  double q[16];
  if (d2pc_oracle_q_from_intrinsics(K1.at<double>(0, 0), K1.at<double>(1, 1), K1.at<double>(0, 2), K1.at<double>(1, 2),
                                    -t.at<double>(0, 0), size.width, size.height, q))
    throw Exception("stereoRectify: bad intrinsics");
  Q.create(4, 4, CV_64FC1);
  for (int i = 0; i < 16; ++i) Q.at<double>(i / 4, i % 4) = q[i];
  (void)R1, (void)R2, (void)P1, (void)P2;  // the reference never reads them
}
}  // namespace cv

// ---------------------------------------------------------------------------------------------------------
// cv_bridge
// ---------------------------------------------------------------------------------------------------------
namespace cv_bridge {
struct Exception : std::runtime_error {
  explicit Exception(const std::string &w) : std::runtime_error(w) {}
};
class CvImage {
 public:
  std_msgs::Header header;
  std::string encoding;
  cv::Mat image;
  void toImageMsg(sensor_msgs::Image &m) const {
    // cv_bridge.cpp CvImage::toImageMsg: dense rows, step = cols * elemSize
    m.header = header;
    m.height = image.rows, m.width = image.cols;
    m.encoding = encoding;
    m.is_bigendian = 0;
    m.step = (uint32_t)((size_t)image.cols * image.elemSize());
    m.data.resize((size_t)m.step * m.height);
    for (int i = 0; i < image.rows; ++i) memcpy(m.data.data() + (size_t)i * m.step, image.data + (size_t)i * image.step, m.step);
  }
  sensor_msgs::ImagePtr toImageMsg() const {
    sensor_msgs::ImagePtr p = std::make_shared<sensor_msgs::Image>();
    toImageMsg(*p);
    return p;
  }
};
typedef boost::shared_ptr<CvImage> CvImagePtr;
inline CvImagePtr toCvCopy(const sensor_msgs::Image &src, const std::string &encoding) {
  if (encoding != "mono8" || (src.encoding != "mono8" && src.encoding != "8UC1"))
    throw Exception("toCvCopy: only mono8 -> mono8 is bound (cv_bridge would convert colour)");
  CvImagePtr p = std::make_shared<CvImage>();
  p->header = src.header;
  p->encoding = encoding;
  p->image.create((int)src.height, (int)src.width, CV_8UC1);
  for (uint32_t i = 0; i < src.height; ++i)
    memcpy(p->image.data + (size_t)i * p->image.step, src.data.data() + (size_t)i * src.step, src.width);
  return p;
}
}  // namespace cv_bridge

// ---------------------------------------------------------------------------------------------------------
// PCL (restated from its public definitions: pcl/impl/point_types.hpp, pcl/point_cloud.h, pcl/conversions.h)
// ---------------------------------------------------------------------------------------------------------
namespace pcl {
struct alignas(16) PointXYZ {
  union {
    float data[4];
    struct {
      float x, y, z;
    };
  };
  PointXYZ() : PointXYZ(0.f, 0.f, 0.f) {}
  PointXYZ(float x_, float y_, float z_) {
    x = x_, y = y_, z = z_;
    data[3] = 1.0f;
  }
};
struct PCLHeader {
  uint32_t seq = 0;
  uint64_t stamp = 0;
  std::string frame_id;
};
template <class PointT>
class PointCloud {
 public:
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  typedef boost::shared_ptr<PointCloud<PointT>> Ptr;
};
inline void toROSMsg(const PointCloud<PointXYZ> &cloud, sensor_msgs::PointCloud2 &msg) {
  // pcl::toPCLPointCloud2 + pcl_conversions::moveFromPCL
  if (cloud.width == 0 && cloud.height == 0) {
    msg.width = (uint32_t)cloud.points.size();
    msg.height = 1;
  } else {
    msg.height = cloud.height;
    msg.width = cloud.width;
  }
  const size_t bytes = sizeof(PointXYZ) * cloud.points.size();
  msg.data.resize(bytes);
  if (bytes) memcpy(msg.data.data(), cloud.points.data(), bytes);
  msg.fields.clear();
  const char *names[3] = {"x", "y", "z"};
  for (int i = 0; i < 3; ++i) {
    sensor_msgs::PointField f;
    f.name = names[i], f.offset = 4u * i, f.datatype = sensor_msgs::PointField::FLOAT32, f.count = 1;
    msg.fields.push_back(f);
  }
  msg.header.seq = cloud.header.seq;
  msg.header.stamp.sec = (uint32_t)(cloud.header.stamp / 1000000ull);
  msg.header.stamp.nsec = (uint32_t)(cloud.header.stamp % 1000000ull) * 1000u;
  msg.header.frame_id = cloud.header.frame_id;
  msg.point_step = sizeof(PointXYZ);
  msg.row_step = (uint32_t)(sizeof(PointXYZ) * msg.width);
  msg.is_dense = cloud.is_dense;
  msg.is_bigendian = 0;
}
}  // namespace pcl
