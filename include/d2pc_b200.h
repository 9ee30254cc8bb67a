/*
 * d2pc_b200.h -- C ABI of libd2pc_b200.so: the B200 (sm_100a) implementation of
 * the per-frame hot path of PX4/disparity_to_point_cloud.
 *
 * The reference has no plugin/operator API: its boundary is the ROS1 callback
 *   void Disparity2PCloud::DisparityCb(const sensor_msgs::ImageConstPtr&)
 *       include/disparity_to_point_cloud/disparity_to_point_cloud.hpp:108
 *       src/disparity_to_point_cloud.cpp:46-92
 * and, for the fusion node, the four callbacks of
 *       include/disparity_to_point_cloud/depth_map_fusion.hpp:127-130
 *       src/depth_map_fusion.cpp:46-136.
 * Each entry point below names the reference lines it replaces.  Plain
 * pointers and sizes only; no C++ or torch types; no exception crosses this
 * boundary; every function returns a d2pc_status (0 = OK, negative = error).
 *
 * Threading (reference: one ros::spin() thread, src/disparity_to_point_cloud_node.cpp:50):
 * a context is single-caller.  Different contexts are independent and may live
 * on different GPUs / host threads.
 *
 * There is NO CPU fallback: every compute entry point fails with
 * D2PC_ERR_NO_DEVICE / D2PC_ERR_CUDA if the GPU path is unavailable.
 *
 * (file:line citations are relative to the reference repository root.)
 */
#ifndef D2PC_B200_H_
#define D2PC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D2PC_ABI_VERSION 1

typedef enum d2pc_status {
  D2PC_OK = 0,
  D2PC_ERR_INVALID_ARG = -1,  /* NULL pointer, bad enum, bad slot ...              */
  D2PC_ERR_BAD_ENCODING = -2, /* image encoding other than mono8 / 8UC1 / 32FC1     */
  D2PC_ERR_BAD_DIMS = -3,     /* width/height/step inconsistent or over capacity    */
  D2PC_ERR_CUDA = -4,         /* a CUDA runtime call failed (see d2pc_last_cuda_error) */
  D2PC_ERR_NO_DEVICE = -5,    /* no usable sm_100 device                            */
  D2PC_ERR_NOMEM = -6,        /* host or device allocation failed                   */
  D2PC_ERR_GEOMETRY = -7,     /* fusion crop rectangle leaves the image (cv::Mat ROI would throw) */
  D2PC_ERR_NOT_READY = -8,    /* wait on a slot with nothing submitted / caches empty */
  D2PC_ERR_BUFFER_TOO_SMALL = -9
} d2pc_status;

/* src/disparity_to_point_cloud.cpp:69-76: the reference's only point filter is
 * the fixed border crop (CROP).  CROP_FINITE is an extension: crop, then drop
 * points with a non-finite x, y or z, order preserved (decoupled look-back
 * stream compaction). */
typedef enum d2pc_filter_mode { D2PC_FILTER_CROP = 0, D2PC_FILTER_CROP_FINITE = 1 } d2pc_filter_mode;

/* EXACT reproduces cv::reprojectImageTo3D's float64 rounding sequence bit for
 * bit (src/disparity_to_point_cloud.cpp:63-64).  FAST is float32 only
 * (<= 1e-5 relative to depth, not bit-exact). */
typedef enum d2pc_arith_mode { D2PC_ARITH_EXACT = 0, D2PC_ARITH_FAST = 1 } d2pc_arith_mode;

/* src/depth_map_fusion.cpp:150-235: the rule getFusedDistance dispatches to.
 * GRAD_FILTER is the one the reference is compiled with. */
typedef enum d2pc_fuse_rule {
  D2PC_FUSE_GRAD_FILTER = 0,
  D2PC_FUSE_MAX_DIST = 1,
  D2PC_FUSE_MAX_DIST_UNLESS_BLACK = 2,
  D2PC_FUSE_BETTER_SCORE = 3,
  D2PC_FUSE_ONLY_GOOD_1 = 4,
  D2PC_FUSE_ONLY_GOOD_AVG = 5,
  D2PC_FUSE_OVERLAP = 6,
  D2PC_FUSE_BLACK_TO_WHITE = 7
} d2pc_fuse_rule;

/* Every constant the reference hard-codes or reads from ROS params, with the
 * reference values as defaults (d2pc_config_default). */
typedef struct d2pc_config {
  uint32_t struct_size; /* sizeof(d2pc_config), for ABI growth */
  /* Disparity2PCloud */
  double fx, fy, cx, cy, baseline; /* hpp:66-71, 84-88: 714.24, 713.5, 376, 240, 0.09 */
  int32_t rect_width, rect_height; /* hpp:101-103: 752 x 480, independent of the frame size */
  int32_t border;                  /* cpp:70,72: 40 */
  int32_t median_ksize;            /* cpp:57: 11 (odd, 1 disables) */
  float disparity_scale;           /* cpp:61: 1/8 */
  int32_t filter_mode;             /* d2pc_filter_mode */
  int32_t arith_mode;              /* d2pc_arith_mode */
  char frame_id[64];               /* cpp:89: "/camera_optical_frame" */
  int32_t verbose;                 /* cpp:82: print "Cloud size: N" when non-zero */
  /* DepthMapFusion */
  int32_t offset_x, offset_y; /* depth_map_fusion.hpp:76-77, launch/depth_map_fusion.launch:8-9 */
  int32_t fuse_rule;          /* d2pc_fuse_rule */
  int32_t fuse_median_ksize;  /* depth_map_fusion.cpp:124: 3 */
  int32_t fuse_crop_left, fuse_crop_right, fuse_crop_top, fuse_crop_bottom; /* :130: 0,40,30,10 */
  /* capacity / pipeline */
  int32_t max_width, max_height; /* largest frame the context will see (buffers are sized once) */
  int32_t max_batch;             /* frames per batched call chunk */
  int32_t n_slots;               /* async pipeline depth (>= 1, default 4) */
} d2pc_config;

typedef struct d2pc_ctx d2pc_ctx;

/* sensor_msgs/PointField as pcl::toROSMsg<pcl::PointXYZ> fills it. */
typedef struct d2pc_point_field {
  char name[8];
  uint32_t offset;
  uint8_t datatype; /* 7 = FLOAT32 */
  uint32_t count;
} d2pc_point_field;

/* The sensor_msgs/PointCloud2 payload DisparityCb publishes
 * (src/disparity_to_point_cloud.cpp:79-90).  `data` is library-owned pinned
 * host memory, valid until the next call that uses the same slot -- or the
 * caller's own buffer when one of the *_into entry points was used. */
typedef struct d2pc_cloud {
  const uint8_t *data; /* width * 16 bytes: {f32 x, f32 y, f32 z, f32 1.0} */
  uint32_t height;     /* 1 */
  uint32_t width;      /* number of points */
  uint32_t point_step; /* 16 */
  uint32_t row_step;   /* 16 * width */
  uint8_t is_bigendian; /* 0 */
  uint8_t is_dense;     /* 0 in CROP mode (cpp:81); 1 in CROP_FINITE */
  uint32_t n_fields;    /* 3 */
  d2pc_point_field fields[3];
} d2pc_cloud;

/* A mono8 sensor_msgs/Image payload (fusion outputs). Library-owned pinned memory. */
typedef struct d2pc_image {
  const uint8_t *data;
  uint32_t width, height, step;
} d2pc_image;

/* ---- lifecycle ------------------------------------------------------------ */

void d2pc_config_default(d2pc_config *cfg);

/* Replaces the Disparity2PCloud / DepthMapFusion constructors
 * (disparity_to_point_cloud.hpp:75-106, depth_map_fusion.hpp:97-124) minus the
 * ROS wiring: selects the device, creates streams, sizes pinned + device
 * buffers and derives Q from the intrinsics in cfg. */
int d2pc_create(const d2pc_config *cfg, int device, d2pc_ctx **out);
void d2pc_destroy(d2pc_ctx *ctx);

/* cv::stereoRectify(K,0,K,0,Size(rect_w,rect_h),I,(-b,0,0)) -> Q, row-major 4x4
 * (disparity_to_point_cloud.hpp:90-104).  Host only. */
int d2pc_q_from_intrinsics(double fx, double fy, double cx, double cy, double baseline, int rect_w, int rect_h,
                           double q_out[16]);
int d2pc_set_q(d2pc_ctx *ctx, const double q[16]);
int d2pc_get_q(const d2pc_ctx *ctx, double q_out[16]);
int d2pc_set_filter_mode(d2pc_ctx *ctx, int filter_mode);
int d2pc_set_arith_mode(d2pc_ctx *ctx, int arith_mode);

/* ---- the disparity callback, host buffers (the reference-facing calls) ------ */

/* Whole DisparityCb (src/disparity_to_point_cloud.cpp:46-92) on one mono8
 * frame in host memory: H2D, median 11, x(1/8), reproject with Q, crop 40,
 * pack {x,y,z,1}, D2H.  Synchronous: `out` is complete on return.  (In CROP mode the kernel of this call stores
 * the points straight into the page-locked destination, so the transfer to the host runs while the kernel does
 * instead of as a copy behind it: tuning key "direct_out".) */
int d2pc_process_mono8(d2pc_ctx *ctx, const uint8_t *data, uint32_t width, uint32_t height, uint32_t step,
                       d2pc_cloud *out);

/* DisparityCb entered after convertTo (cpp:63 onwards) on a float disparity
 * frame in host memory; step in bytes. */
int d2pc_process_f32(d2pc_ctx *ctx, const float *disp, uint32_t width, uint32_t height, uint32_t step,
                     d2pc_cloud *out);

/* Asynchronous variants on pipeline slot `slot` (0 <= slot < n_slots): H2D,
 * kernels and D2H are enqueued on three streams chained by events, so slots
 * overlap.  The input buffer may be reused once d2pc_wait(slot) returns (it is
 * read by the DMA engine directly when it is pinned -- see d2pc_host_alloc --
 * and staged through library pinned memory otherwise, in which case it may be
 * reused as soon as submit returns). */
int d2pc_submit_mono8(d2pc_ctx *ctx, int slot, const uint8_t *data, uint32_t width, uint32_t height, uint32_t step);
int d2pc_submit_f32(d2pc_ctx *ctx, int slot, const float *disp, uint32_t width, uint32_t height, uint32_t step);
int d2pc_wait(d2pc_ctx *ctx, int slot, d2pc_cloud *out);

/* The same calls with a CALLER-SUPPLIED destination for the cloud (SURVEY.md 8(b) "Ownership"): what
 * pcl::toROSMsg's memcpy into PointCloud2.data does in the reference (src/disparity_to_point_cloud.cpp:84-85)
 * becomes the device-to-host copy itself.  `dst` must hold the whole cloud (16 bytes per point; `cap` bytes are
 * available, D2PC_ERR_BUFFER_TOO_SMALL otherwise -- in CROP mode at submit, in CROP_FINITE mode, where the count
 * is only known afterwards, at wait).  When `dst` is page-locked (d2pc_host_alloc, or any buffer passed through
 * d2pc_host_register, e.g. the storage of a message's data vector) and 16-byte aligned the DMA engine writes
 * straight into it; a pageable `dst` is filled from the library's pinned buffer when the slot is waited on.
 * `dst` must stay valid until d2pc_wait(slot) returns; the cloud returned by d2pc_wait then points at `dst`. */
int d2pc_submit_mono8_into(d2pc_ctx *ctx, int slot, const uint8_t *data, uint32_t width, uint32_t height,
                           uint32_t step, uint8_t *dst, size_t cap);
int d2pc_submit_f32_into(d2pc_ctx *ctx, int slot, const float *disp, uint32_t width, uint32_t height, uint32_t step,
                         uint8_t *dst, size_t cap);
int d2pc_process_mono8_into(d2pc_ctx *ctx, const uint8_t *data, uint32_t width, uint32_t height, uint32_t step,
                            uint8_t *dst, size_t cap, d2pc_cloud *out);
int d2pc_process_f32_into(d2pc_ctx *ctx, const float *disp, uint32_t width, uint32_t height, uint32_t step,
                          uint8_t *dst, size_t cap, d2pc_cloud *out);

/* Per-call timing (the reference's only instrumentation is nine printf lines per frame,
 * src/disparity_to_point_cloud.cpp:47-91).  d2pc_set_timing(ctx, 1) makes the slots' events carry timestamps
 * (waits for the device; call it between frames); after d2pc_wait(slot) d2pc_slot_timing reports where that
 * submission's time went on the device: the spans between the events that chain H2D copy -> kernels -> D2H copy
 * (queueing behind other slots included).  In CROP_FINITE mode d2h_us covers the 4-byte count only (the payload
 * is copied inside d2pc_wait); where the kernel stores the cloud into host memory itself ("direct_out": the
 * synchronous mono8 entries) the transfer is part of kernels_us and d2h_us is ~0; a synchronous call into a pageable
 * caller buffer moves its cloud inside d2pc_wait, outside these spans.  The host-side spans of every entry point are also NVTX ranges ("d2pc submit ...",
 * "d2pc H2D", "d2pc kernels", "d2pc D2H") for Nsight. */
typedef struct d2pc_timing {
  float h2d_us, kernels_us, d2h_us, total_us;
  uint64_t points; /* points of that cloud */
} d2pc_timing;
int d2pc_set_timing(d2pc_ctx *ctx, int enable);
int d2pc_slot_timing(d2pc_ctx *ctx, int slot, d2pc_timing *out);

/* Pinned host memory for buffers the caller wants DMA'd without staging: allocate it here, or page-lock memory
 * the caller already owns (cudaHostRegister; costs about as much as touching every page once, so register
 * long-lived buffers, not one per message). */
int d2pc_host_alloc(void **ptr, size_t bytes);
int d2pc_host_free(void *ptr);
int d2pc_host_register(void *ptr, size_t bytes);
int d2pc_host_unregister(void *ptr);

/* Streams `n_frames` same-sized frames through the slot pipeline and hands every finished cloud to `sink` in
 * frame order (sink may be NULL: the clouds are then produced and dropped).  Frame i is read from
 * frames + (i % ring_len) * frame_stride bytes (ring_len 0 means n_frames, i.e. a plain array); is_f32 selects
 * the float vs mono8 entry.  This is the loop a subscriber thread would run; it is what bench.py's end-to-end
 * figure times. */
typedef void (*d2pc_cloud_sink)(void *user, uint64_t frame_index, const d2pc_cloud *cloud);
int d2pc_process_stream(d2pc_ctx *ctx, const void *frames, uint64_t n_frames, size_t frame_stride, uint64_t ring_len,
                        int is_f32, uint32_t width, uint32_t height, uint32_t step, d2pc_cloud_sink sink,
                        void *user);

/* ---- device-resident entry points (kernel-only figures, chaining) ---------- */

/* All pointers are DEVICE pointers on the context's device; work is enqueued
 * on the context's compute stream and NOT synchronised (d2pc_sync, or CUDA
 * events recorded on d2pc_compute_stream()). */

/* cpp:63-85 for a batch of float frames: frame f at d_disp + f*frame_stride
 * bytes, rows `step` bytes apart.  Points of frame f start at
 * d_points + f*points_stride bytes.  In CROP mode every frame produces
 * (w-2b)*(h-2b) points; in CROP_FINITE mode d_counts[f] (uint32, may be NULL
 * in CROP mode) receives the number of points kept. */
int d2pc_reproject_f32_device(d2pc_ctx *ctx, const float *d_disp, uint32_t n_frames, uint32_t width,
                              uint32_t height, size_t step, size_t frame_stride, uint8_t *d_points,
                              size_t points_stride, uint32_t *d_counts);

/* cpp:55-85 for a batch of mono8 frames (median ksize from the config). */
int d2pc_reproject_mono8_device(d2pc_ctx *ctx, const uint8_t *d_img, uint32_t n_frames, uint32_t width,
                                uint32_t height, size_t step, size_t frame_stride, uint8_t *d_points,
                                size_t points_stride, uint32_t *d_counts);

/* cv::medianBlur on CV_8UC1, replicate border (cpp:55-57 with ksize 11,
 * depth_map_fusion.cpp:124 with ksize 3).  ksize odd, 3..15. */
int d2pc_median_u8_device(d2pc_ctx *ctx, const uint8_t *d_src, uint32_t width, uint32_t height, size_t src_step,
                          uint8_t *d_dst, size_t dst_step, int ksize);

void *d2pc_compute_stream(d2pc_ctx *ctx); /* cudaStream_t */
int d2pc_sync(d2pc_ctx *ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t d2pc_launch_count(const d2pc_ctx *ctx);

/* ---- depth_map_fusion -------------------------------------------------------- */

/* Geometry of src/depth_map_fusion.cpp:237-265 for a w x h frame: rect1 (map /
 * score 1), rect2 (map / score 2, in the rotated frame), rectc (output
 * container), each {x, y, w, h}; dims = {n, fused_w, fused_h}. */
int d2pc_fuse_geometry(const d2pc_ctx *ctx, uint32_t width, uint32_t height, int rect1[4], int rect2[4],
                       int rectc[4], int dims[3]);

/* One DisparityCb1 + DisparityCb2 + publishFusedDepthMap pass
 * (src/depth_map_fusion.cpp:46-62, 103-136) on four same-sized mono8 frames in
 * host memory: rotate map/score 2, crop all four to the common square, merge
 * per pixel, median 3, final border trim.  s1/s2 are the preprocessed score
 * images (what the caches cropped_score_{1,2}_ hold, :77/:96).
 * fused  -> what is published on /fused_depth_map (:134-136)
 * combined -> what is published on /combined_score (:126-127); may be NULL. */
int d2pc_fuse(d2pc_ctx *ctx, const uint8_t *d1, const uint8_t *d2, const uint8_t *s1, const uint8_t *s2,
              uint32_t width, uint32_t height, uint32_t step, d2pc_image *fused, d2pc_image *combined);

/* MatchingScoreCb1 (which = 1, src/depth_map_fusion.cpp:64-80) / MatchingScoreCb2 (which = 2, :82-99) on one
 * mono8 score frame in host memory: [rotate for 2 ->] cropToSquare -> GaussianBlur 13 sigma 3 -> Sobel 2nd
 * derivative ksize 7 x0.03 -> threshold 30 -> GaussianBlur 21 sigma 10 -> score + 2*grad (saturating).
 * out is the n x n image the node caches as cropped_score_k_ (library-owned pinned memory, one buffer per
 * `which`, valid until the next call with the same `which`). */
int d2pc_preprocess_score(d2pc_ctx *ctx, const uint8_t *score, uint32_t width, uint32_t height, uint32_t step,
                          int which, d2pc_image *out);
/* Same with device pointers; d_out is n x n dense. */
int d2pc_preprocess_score_device(d2pc_ctx *ctx, const uint8_t *d_score, uint32_t width, uint32_t height, size_t step,
                                 int which, uint8_t *d_out);

/* d2pc_fuse with the two score inputs given as the n x n preprocessed caches (d2pc_preprocess_score output)
 * instead of full frames: exactly the state publishFusedDepthMap works from. */
int d2pc_fuse_preprocessed(d2pc_ctx *ctx, const uint8_t *d1, const uint8_t *d2, const uint8_t *s1_cropped,
                           const uint8_t *s2_cropped, uint32_t width, uint32_t height, uint32_t step,
                           d2pc_image *fused, d2pc_image *combined);

/* Same, device pointers, enqueued on the compute stream.  d_fused is
 * fused_w x fused_h dense; d_combined n x n dense (may be NULL). */
int d2pc_fuse_device(d2pc_ctx *ctx, const uint8_t *d_d1, const uint8_t *d_d2, const uint8_t *d_s1,
                     const uint8_t *d_s2, uint32_t width, uint32_t height, size_t step, uint8_t *d_fused,
                     uint8_t *d_combined);

/* d2pc_fuse_device with the two scores given as the n x n preprocessed caches (d2pc_preprocess_score_device
 * output, dense): the device-resident form of d2pc_fuse_preprocessed. */
int d2pc_fuse_preprocessed_device(d2pc_ctx *ctx, const uint8_t *d_d1, const uint8_t *d_d2, const uint8_t *d_s1_cropped,
                                  const uint8_t *d_s2_cropped, uint32_t width, uint32_t height, size_t step,
                                  uint8_t *d_fused, uint8_t *d_combined);

/* DepthMapFusion::colorizeDepth (src/depth_map_fusion.cpp:304-358), the RAINBOW_WITH_BLACK colouring of the
 * node's debug views: mono8 in host memory -> 3 bytes per pixel in the byte order the reference stores (and labels
 * "rgb8", :298-299).  out->width = width, out->step = 3 * width. */
int d2pc_colorize_depth(d2pc_ctx *ctx, const uint8_t *gray, uint32_t width, uint32_t height, uint32_t step,
                        d2pc_image *out);

/* BASELINE config 5: fuse four host frames, then run the whole DisparityCb on
 * the fused map without leaving the device. */
int d2pc_fuse_then_process(d2pc_ctx *ctx, const uint8_t *d1, const uint8_t *d2, const uint8_t *s1,
                           const uint8_t *s2, uint32_t width, uint32_t height, uint32_t step, d2pc_cloud *out);

/* The same frame set on the slot pipeline (asynchronous, like d2pc_submit_mono8): H2D of the four frames,
 * [MatchingScoreCb1/2 when preprocess_scores != 0: s1 / s2 are then the raw matching-score frames of
 * src/depth_map_fusion.cpp:64-99; with 0 they are full frames that already hold the preprocessed scores, as for
 * d2pc_fuse], the merge, median 3 + trim, DisparityCb on the fused map and the D2H of the cloud are enqueued on the
 * three streams; collect the cloud with d2pc_wait(ctx, slot, &cloud).  Slots overlap frame sets. */
int d2pc_submit_fusion(d2pc_ctx *ctx, int slot, const uint8_t *d1, const uint8_t *d2, const uint8_t *s1,
                       const uint8_t *s2, uint32_t width, uint32_t height, uint32_t step, int preprocess_scores);

/* Streams n_sets frame sets through d2pc_submit_fusion / d2pc_wait with every slot busy; set i is read from
 * sets + (i % ring_len) * set_stride (ring_len 0 = n_sets) as four frames of step * height bytes back to back in
 * the order d1, d2, s1, s2.  What bench.py --config 5 times end to end. */
int d2pc_process_fusion_stream(d2pc_ctx *ctx, const uint8_t *sets, uint64_t n_sets, size_t set_stride,
                               uint64_t ring_len, uint32_t width, uint32_t height, uint32_t step,
                               int preprocess_scores, d2pc_cloud_sink sink, void *user);

/* ---- ROS1 wire helpers (what Publisher::publish serialises) ------------------- */

/* sensor_msgs/PointCloud2 with header {seq, stamp, cfg.frame_id}; returns the
 * byte count (write happens only if cap suffices; pass NULL/0 to size). */
size_t d2pc_serialize_pointcloud2(const d2pc_ctx *ctx, const d2pc_cloud *cloud, uint32_t seq, uint32_t stamp_sec,
                                  uint32_t stamp_nsec, uint8_t *out, size_t cap);

/* ---- diagnostics ---------------------------------------------------------------- */

/* Tuning / test hook: integer knobs by name.  Not needed in normal use.
 *   kernels : "rows_per_unit" (CROP: rows per work unit; band / pipeline kernels: rows per tile), "ctas_per_sm",
 *             "median_strip", "median_variant" (0 default: window histogram, 3x3 selection network; 2 window
 *             histogram for every size),
 *             "compact_variant" (0 auto: band kernel where Q allows; 1 park kernel for every Q),
 *             "prefetch_dist" (L2 prefetch distance in work units / tiles; 0 automatic, < 0 off),
 *             "exact_variant" (0 guarded multiply, 1 Markstein), "zero_numer" (kernel variant that keeps an
 *             exactly-zero X numerator column straight-line: 0 when Q has such a column, 1 always, -1 never),
 *             "fuse_median" (mono8 callback: 0 one fused median + reproject launch where Q allows, -1 always two),
 *             "direct_out" (CROP clouds stored by the kernel straight into the page-locked destination instead of
 *             a D2H copy behind the kernel: 0 the synchronous mono8 entries only, 1 every submission, -1 never),
 *             "force_scalar", "force_generic"
 *   config  : "median_ksize", "border", "offset_x", "offset_y", "fuse_rule", "fuse_median_ksize",
 *             "fuse_crop_left|right|top|bottom" */
int d2pc_set_tuning(d2pc_ctx *ctx, const char *key, int value);

const char *d2pc_strerror(int status);
const char *d2pc_last_cuda_error(const d2pc_ctx *ctx);
int d2pc_abi_version(void);
/* Devices with compute capability 10.x visible to the process. */
int d2pc_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* D2PC_B200_H_ */
