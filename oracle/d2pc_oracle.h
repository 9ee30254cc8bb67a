/*
 * d2pc_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A dependency-free plain-C restatement of the per-frame hot path of
 * PX4/disparity_to_point_cloud.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library, and
 * only as the checker / the CPU baseline -- never on the product path.
 *
 * Parity pin: the reference has NO tests, fixtures or golden vectors
 * (CMakeLists.txt:206-217 has the gtest stanza commented out) and its own build
 * system cannot run here (needs ROS1, cv_bridge, OpenCV C++ and PCL, none
 * installed).  This oracle is pinned two ways:
 *   * third-party ARITHMETIC (un-vendored, un-pinned libraries, package.xml:41-55)
 *     against Python cv2 4.13.0 -- the same OpenCV entry points the reference
 *     calls, driven with the reference's arguments: tests/golden/make_golden.py
 *     generates the committed fixtures, tests/test_oracle_*.py replay them;
 *   * the reference's OWN LOGIC against the reference's own code: oracle/_ref is
 *     src/depth_map_fusion.cpp + src/disparity_to_point_cloud.cpp compiled
 *     unmodified against stand-in headers (oracle/ref_stubs/), and
 *     tests/test_ref_compiled.py compares this file with it (gradFilter
 *     exhaustively, the seven rules, the crop geometry, colorizeDepth, callback
 *     sequences, DisparityCb's cloud).
 * PCL's struct layout / toROSMsg, cv_bridge's copy and the ROS1 wire format
 * are restated from their public definitions.
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef D2PC_ORACLE_H_
#define D2PC_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* include/disparity_to_point_cloud/disparity_to_point_cloud.hpp:90-104
 * cv::stereoRectify(K, 0, K, 0, Size(rect_w, rect_h), I, (-baseline,0,0)) -> Q
 * restated for that degenerate call (zero distortion, R = I, horizontal rig,
 * CALIB_ZERO_DISPARITY, alpha = -1). q is row-major 4x4. */
int d2pc_oracle_q_from_intrinsics(double fx, double fy, double cx, double cy,
                                  double baseline, int rect_w, int rect_h,
                                  double q[16]);

/* src/disparity_to_point_cloud.cpp:55-57 and src/depth_map_fusion.cpp:124
 * cv::medianBlur(src, dst, ksize) on CV_8UC1, replicate border. ksize odd.
 * dst must not alias src. */
void d2pc_oracle_median_blur_u8(const uint8_t *src, int w, int h,
                                size_t src_step, uint8_t *dst, size_t dst_step,
                                int ksize);

/* src/disparity_to_point_cloud.cpp:60-61  Mat::convertTo(CV_32FC1, alpha) */
void d2pc_oracle_convert_u8_f32(const uint8_t *src, int w, int h,
                                size_t src_step, float *dst, size_t dst_step,
                                double alpha);

/* src/disparity_to_point_cloud.cpp:63-64
 * cv::reprojectImageTo3D(disp, out, Q, handleMissingValues=false), CV_32F in,
 * CV_32FC3 out (dense, 3*w floats per row).  Rounding per SURVEY.md A.2. */
void d2pc_oracle_reproject_image_to_3d(const float *disp, int w, int h,
                                       size_t disp_step, const double q[16],
                                       float *xyz);

/* src/disparity_to_point_cloud.cpp:69-85: border crop loop with push_back of
 * pcl::PointXYZ {x,y,z,1.0f} into an un-reserved growing vector, then the
 * pcl::toROSMsg memcpy into cloud.  Returns the point count
 * max(0,w-2*border)*max(0,h-2*border).  cloud may be NULL to query the count. */
size_t d2pc_oracle_crop_pack(const float *xyz, int w, int h, int border,
                             uint8_t *cloud);

/* src/disparity_to_point_cloud.cpp:46-92, all stages with the reference's
 * intermediate buffers: mono8 copy -> median 11 -> x(1/8) -> reproject ->
 * crop 40 -> pack.  Returns the point count; cloud holds count*16 bytes. */
size_t d2pc_oracle_disparity_cb_mono8(const uint8_t *img, int w, int h,
                                      size_t step, const double q[16],
                                      uint8_t *cloud);

/* The same callback entered after convertTo (:63 onwards): the float-entry
 * path BASELINE.json's configs 1, 3 and 4 are quoted on. */
size_t d2pc_oracle_disparity_cb_f32(const float *disp, int w, int h,
                                    size_t step, const double q[16],
                                    uint8_t *cloud);

/* Extension mode CROP_FINITE: the crop output filtered
 * by isfinite(x) && isfinite(y) && isfinite(z), order preserved. */
size_t d2pc_oracle_filter_finite(const uint8_t *cloud, size_t n_points,
                                 uint8_t *out);

/* ROS1 wire image of the sensor_msgs/PointCloud2 the node publishes
 * (src/disparity_to_point_cloud.cpp:79-90; layout SURVEY.md A.3).
 * Returns bytes needed; writes only if cap is large enough. */
size_t d2pc_oracle_serialize_pointcloud2(uint32_t seq, uint32_t sec,
                                         uint32_t nsec, const char *frame_id,
                                         const uint8_t *points,
                                         uint32_t n_points, uint8_t is_dense,
                                         uint8_t *out, size_t cap);

/* ---- depth_map_fusion ---- */

/* src/depth_map_fusion.cpp:247-265 cropToSquare; member_offset_y is the
 * class member offset_y_ the function reads instead of its parameter (:252).
 * rect = {x, y, w, h}. Returns 0, or -1 if the rectangle leaves the image
 * (where cv::Mat::operator() would throw). */
int d2pc_oracle_crop_to_square(int cols, int rows, int offset_x, int offset_y,
                               int member_offset_y, int rect[4]);

/* src/depth_map_fusion.cpp:268-273 rotateMat: 90 degrees clockwise.
 * src is w x h, dst is h x w (cols = h), dense. */
void d2pc_oracle_rotate_cw(const uint8_t *src, int w, int h, size_t src_step,
                           uint8_t *dst);

/* src/depth_map_fusion.cpp:219-235 gradFilter. */
int d2pc_oracle_grad_filter(int dist1, int dist2, int score1, int score2,
                            int grad1, int grad2);

/* gradFilter over all (dist1, dist2) at fixed scores; out[dist1*256 + dist2] (test helper). */
void d2pc_oracle_grad_filter_table(int score1, int score2, uint8_t *out);

/* src/depth_map_fusion.cpp:162-217, the alternate (unused) fusion rules,
 * selected by mode: 0 gradFilter, 1 maxDist, 2 maxDistUnlessBlack,
 * 3 betterScore, 4 onlyGood1, 5 onlyGoodAvg, 6 overlap, 7 blackToWhite. */
int d2pc_oracle_fuse_rule(int mode, int dist1, int dist2, int score1,
                          int score2);

/* One pass of DisparityCb1 + DisparityCb2 + publishFusedDepthMap
 * (src/depth_map_fusion.cpp:46-62, 103-136) on four same-sized mono8 images:
 * d1, s1 cropped with (+ox,+oy); d2, s2 rotated then cropped with (-ox,-oy);
 * the merge loop; medianBlur 3 on the output container (the un-rotated d2
 * frame cropped with (0,0)); cropMat(0,40,30,10).
 * s1/s2 are the already-preprocessed score images (the caches
 * cropped_score_{1,2}_ == cropped_score_{1,2}_grad_, :77, :96) at full frame
 * size, i.e. score preprocessing is outside this function.
 * fused: out_w x out_h bytes (dense); combined: n x n bytes (dense) -- the
 * aliased cropped_score_combined_/cropped_score_1_ buffer after the loop.
 * dims = {n, out_w, out_h}.  Returns 0, -1 bad geometry. */
int d2pc_oracle_fuse(const uint8_t *d1, const uint8_t *d2, const uint8_t *s1,
                     const uint8_t *s2, int w, int h, size_t step,
                     int offset_x, int offset_y, int mode, uint8_t *fused,
                     uint8_t *combined, int dims[3]);

/* src/depth_map_fusion.cpp:64-80 (vertical = 0, MatchingScoreCb1) / :82-99 (vertical = 1, MatchingScoreCb2):
 * what the node caches as cropped_score_k_ (== cropped_score_k_grad_) for a score frame (already rotated for
 * callback 2) and its cropToSquare rectangle rect = {x, y, n, n}:
 *   GaussianBlur 13x13 sigma 3 (the crop is a non-isolated submatrix: the blur sees the frame around it and, being
 *   a submatrix, runs as sepFilter2D with float32 kernels, not OpenCV's fixed-point path) ->
 *   Sobel 2nd derivative ksize 7 scale 0.03 -> threshold 30 -> GaussianBlur 21x21 sigma 10 (stand-alone Mat: the
 *   fixed-point path) -> score + 2*grad.
 * Arithmetic pinned against cv2 4.13.0 (AVX2 dispatch): float32 separable filters whose vector loops use FMA and
 * whose scalar tail columns do not; 8.8 fixed-point kernel for the second blur.  out is n x n dense.
 * Returns 0, -1 on bad geometry. */
int d2pc_oracle_score_preprocess(const uint8_t *frame, int w, int h, size_t step, const int rect[4], int vertical,
                                 uint8_t *out);

/* cv::GaussianBlur(parent(rect), dst, Size(ksize, ksize), sigma) on CV_8U for the two calls the reference makes
 * (13 / 3.0 at depth_map_fusion.cpp:70-71, :89-90 and 21 / 10.0 at :74-75, :93-94).  submatrix = Mat::isSubmatrix()
 * of the source: OpenCV 4.x runs its fixed-point path only for non-submatrix (or isolated-border) sources and
 * sepFilter2D with float32 kernels otherwise.  Returns 0, -1 bad rectangle, -2 unsupported kernel. */
int d2pc_oracle_gaussian_blur_u8(const uint8_t *parent, int pw, int ph, size_t step, const int rect[4], int ksize,
                                 double sigma, int submatrix, uint8_t *dst, size_t dst_step);

/* cv::Sobel(src, dst, -1, dx, dy, 7, 0.03), (dx, dy) = (0, 2) [vertical = 0] / (2, 0) [vertical = 1], n x n dense. */
void d2pc_oracle_sobel7_second_u8(const uint8_t *src, int n, int vertical, uint8_t *dst);

/* src/depth_map_fusion.cpp:304-358 colorizeDepth (the RAINBOW_WITH_BLACK debug views); rgb is w*h*3 dense. */
void d2pc_oracle_colorize_depth(const uint8_t *gray, int w, int h, size_t step, uint8_t *rgb);

/* ---- CPU baseline helpers (bench.py cpu_baseline / --impl reference) ---- */

/* Runs d2pc_oracle_disparity_cb_f32 (mono8 == 0) or _mono8 (mono8 != 0) over
 * n_frames frames laid out back to back (frame stride = step*h bytes), frame
 * parallel on n_threads pthreads; every frame writes its cloud to
 * cloud + i*cloud_stride.  Returns total points. */
size_t d2pc_oracle_run_frames(const void *frames, int n_frames, int w, int h,
                              size_t step, int mono8, const double q[16],
                              uint8_t *cloud, size_t cloud_stride,
                              int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* D2PC_ORACLE_H_ */
