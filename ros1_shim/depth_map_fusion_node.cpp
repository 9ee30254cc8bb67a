// ROS1 shim: the reference's depth_map_fusion_node with the per-pixel merge on the GPU (libd2pc_b200.so).
// Built only where catkin/roscpp exist; NOT built in the development image.
// Subscriptions, the /fused_depth_map publisher, the offset_x / offset_y parameters follow
// include/disparity_to_point_cloud/depth_map_fusion.hpp:97-124, so launch/depth_map_fusion.launch works unchanged.
// Score frames are preprocessed on the GPU as they arrive (Gaussian / Sobel / threshold / Gaussian chain of
// src/depth_map_fusion.cpp:64-99) and cached as the n x n images the reference caches; debug views are not
// published.
#include <ros/ros.h>
#include <sensor_msgs/Image.h>
#include <sensor_msgs/image_encodings.h>

#include <cstring>
#include <vector>

#include "d2pc_b200.h"

namespace {

class DepthMapFusionGpu {
 public:
  DepthMapFusionGpu() : nh_("~") {
    d1_sub_ = nh_.subscribe("/disparity_1", 1, &DepthMapFusionGpu::DisparityCb1, this);
    d2_sub_ = nh_.subscribe("/disparity_2", 1, &DepthMapFusionGpu::DisparityCb2, this);
    s1_sub_ = nh_.subscribe("/matching_score_1", 1, &DepthMapFusionGpu::MatchingScoreCb1, this);
    s2_sub_ = nh_.subscribe("/matching_score_2", 1, &DepthMapFusionGpu::MatchingScoreCb2, this);
    fused_pub_ = nh_.advertise<sensor_msgs::Image>("/fused_depth_map", 5);
    d2pc_config cfg;
    d2pc_config_default(&cfg);
    if (!nh_.getParam("offset_x", cfg.offset_x)) ROS_WARN("Failed to load parameter offset_x");
    if (!nh_.getParam("offset_y", cfg.offset_y)) ROS_WARN("Failed to load parameter offset_y");
    const int rc = d2pc_create(&cfg, 0, &ctx_);
    if (rc != D2PC_OK) {
      ROS_FATAL("d2pc_create: %s", d2pc_strerror(rc));
      ros::shutdown();
    }
  }
  ~DepthMapFusionGpu() { d2pc_destroy(ctx_); }

  void DisparityCb1(const sensor_msgs::ImageConstPtr &m) { d1_ = m; }
  void MatchingScoreCb1(const sensor_msgs::ImageConstPtr &m) { Preprocess(m, 1, s1_); }
  void MatchingScoreCb2(const sensor_msgs::ImageConstPtr &m) { Preprocess(m, 2, s2_); }
  void Preprocess(const sensor_msgs::ImageConstPtr &m, int which, std::vector<uint8_t> &slot) {
    d2pc_image out;
    if (d2pc_preprocess_score(ctx_, m->data.data(), m->width, m->height, m->step, which, &out) != D2PC_OK) return;
    slot.assign(out.data, out.data + static_cast<size_t>(out.step) * out.height);
  }
  void DisparityCb2(const sensor_msgs::ImageConstPtr &m) {
    d2_ = m;
    if (!d1_ || s1_.empty() || s2_.empty()) return;  // have not received all maps and scores yet
    if (d1_->width != m->width || d1_->height != m->height || d1_->step != m->step) return;
    d2pc_image fused, combined;
    const int rc = d2pc_fuse_preprocessed(ctx_, d1_->data.data(), d2_->data.data(), s1_.data(), s2_.data(), m->width,
                                          m->height, m->step, &fused, &combined);
    if (rc != D2PC_OK) {
      ROS_ERROR("d2pc_fuse_preprocessed: %s", d2pc_strerror(rc));
      return;
    }
    // the reference's combined score aliases its cached score 1 (src/depth_map_fusion.cpp:113)
    s1_.assign(combined.data, combined.data + static_cast<size_t>(combined.step) * combined.height);
    sensor_msgs::Image out;
    out.header = m->header;
    out.height = fused.height;
    out.width = fused.width;
    out.encoding = sensor_msgs::image_encodings::MONO8;
    out.is_bigendian = 0;
    out.step = fused.width;
    out.data.resize(static_cast<size_t>(fused.width) * fused.height);
    for (uint32_t y = 0; y < fused.height; ++y)
      std::memcpy(&out.data[static_cast<size_t>(y) * fused.width], fused.data + static_cast<size_t>(y) * fused.step,
                  fused.width);
    fused_pub_.publish(out);
  }

 private:
  ros::NodeHandle nh_;
  ros::Subscriber d1_sub_, d2_sub_, s1_sub_, s2_sub_;
  ros::Publisher fused_pub_;
  sensor_msgs::ImageConstPtr d1_, d2_;
  std::vector<uint8_t> s1_, s2_;
  d2pc_ctx *ctx_ = nullptr;
};

}  // namespace

int main(int argc, char **argv) {
  ros::init(argc, argv, "depth_map_fusion");
  DepthMapFusionGpu node;
  ros::spin();
  return 0;
}
