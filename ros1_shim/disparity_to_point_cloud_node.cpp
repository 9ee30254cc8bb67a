// ROS1 shim: the reference's disparity_to_point_cloud_node with DisparityCb's body replaced by libd2pc_b200.so.
// Built only where catkin/roscpp exist (see ros1_shim/CMakeLists.txt); NOT built in the development image.
// Topic names, queue sizes, the (accidentally) latched publisher, parameter names and defaults follow
// include/disparity_to_point_cloud/disparity_to_point_cloud.hpp:75-106 of the reference, so
// launch/d2pcloud.launch works unchanged.
#include <ros/ros.h>
#include <sensor_msgs/Image.h>
#include <sensor_msgs/PointCloud2.h>
#include <sensor_msgs/image_encodings.h>

#include <cstring>

#include "d2pc_b200.h"

namespace {

class Disparity2PCloudGpu {
 public:
  Disparity2PCloudGpu() : nh_("~") {
    sub_ = nh_.subscribe("/disparity", 1, &Disparity2PCloudGpu::DisparityCb, this);
    pub_ = nh_.advertise<sensor_msgs::PointCloud2>("/point_cloud", 1, /*latch=*/true);
    d2pc_config cfg;
    d2pc_config_default(&cfg);
    nh_.param<double>("fx_", cfg.fx, 714.24);
    nh_.param<double>("fy_", cfg.fy, 713.5);
    nh_.param<double>("cx_", cfg.cx, 376);
    nh_.param<double>("cy_", cfg.cy, 240);
    nh_.param<double>("base_line_", cfg.baseline, 0.09);
    int device = 0;
    nh_.param<int>("cuda_device", device, 0);
    const int rc = d2pc_create(&cfg, device, &ctx_);
    if (rc != D2PC_OK) {
      ROS_FATAL("d2pc_create: %s", d2pc_strerror(rc));
      ros::shutdown();
    }
  }
  ~Disparity2PCloudGpu() { d2pc_destroy(ctx_); }

  void DisparityCb(const sensor_msgs::ImageConstPtr &msg) {
    namespace enc = sensor_msgs::image_encodings;
    if (msg->encoding != enc::MONO8 && msg->encoding != enc::TYPE_8UC1) {
      ROS_ERROR_THROTTLE(1.0, "unsupported encoding '%s' (mono8 only)", msg->encoding.c_str());
      return;
    }
    d2pc_cloud cloud;
    const int rc = d2pc_process_mono8(ctx_, msg->data.data(), msg->width, msg->height, msg->step, &cloud);
    if (rc != D2PC_OK) {
      ROS_ERROR("d2pc_process_mono8: %s (%s)", d2pc_strerror(rc), d2pc_last_cuda_error(ctx_));
      return;
    }
    sensor_msgs::PointCloud2 out;
    out.height = cloud.height;
    out.width = cloud.width;
    out.fields.resize(cloud.n_fields);
    for (uint32_t i = 0; i < cloud.n_fields; ++i) {
      out.fields[i].name = cloud.fields[i].name;
      out.fields[i].offset = cloud.fields[i].offset;
      out.fields[i].datatype = cloud.fields[i].datatype;
      out.fields[i].count = cloud.fields[i].count;
    }
    out.is_bigendian = cloud.is_bigendian;
    out.point_step = cloud.point_step;
    out.row_step = cloud.row_step;
    out.is_dense = cloud.is_dense;
    out.data.resize(static_cast<size_t>(cloud.row_step) * cloud.height);
    std::memcpy(out.data.data(), cloud.data, out.data.size());
    out.header.stamp = msg->header.stamp;
    out.header.frame_id = "/camera_optical_frame";
    pub_.publish(out);
  }

 private:
  ros::NodeHandle nh_;
  ros::Subscriber sub_;
  ros::Publisher pub_;
  d2pc_ctx *ctx_ = nullptr;
};

}  // namespace

int main(int argc, char **argv) {
  ros::init(argc, argv, "disparity_to_point_cloud");
  Disparity2PCloudGpu node;
  ros::spin();
  return 0;
}
