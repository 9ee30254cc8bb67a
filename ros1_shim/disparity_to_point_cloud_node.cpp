// ROS1 shim: the reference's disparity_to_point_cloud_node with DisparityCb's body replaced by libd2pc_b200.so.
// Built with catkin where roscpp exists (ros1_shim/package.xml, CMakeLists.txt); in the development image it is
// compiled and linked against the API stand-ins of tests/ros_stubs/ (tests/test_ros1_shim.py).
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
    border_ = cfg.border;
    const int rc = d2pc_create(&cfg, device, &ctx_);
    if (rc != D2PC_OK) {
      ROS_FATAL("d2pc_create: %s", d2pc_strerror(rc));
      ros::shutdown();
    }
  }
  ~Disparity2PCloudGpu() {
    d2pc_destroy(ctx_);
    if (registered_) d2pc_host_unregister(registered_);
  }

  void DisparityCb(const sensor_msgs::ImageConstPtr &msg) {
    namespace enc = sensor_msgs::image_encodings;
    if (msg->encoding != enc::MONO8 && msg->encoding != enc::TYPE_8UC1) {
      ROS_ERROR_THROTTLE(1.0, "unsupported encoding '%s' (mono8 only)", msg->encoding.c_str());
      return;
    }
    // the cloud is DMA'd straight into the message's (page-locked) data vector: no counterpart of the
    // pcl::toROSMsg memcpy (reference cpp:84-85)
    const long cw = static_cast<long>(msg->width) - 2L * border_, ch = static_cast<long>(msg->height) - 2L * border_;
    const size_t bytes = cw > 0 && ch > 0 ? static_cast<size_t>(cw) * static_cast<size_t>(ch) * 16 : 0;
    sensor_msgs::PointCloud2 &out = out_;
    if (out.data.capacity() < bytes || out.data.data() != registered_) {
      if (registered_) d2pc_host_unregister(registered_);
      registered_ = nullptr;
      out.data.clear();
      out.data.reserve(bytes + bytes / 4 + 4096);
      out.data.resize(bytes);
      if (d2pc_host_register(out.data.data(), out.data.capacity()) == D2PC_OK) registered_ = out.data.data();
    } else if (out.data.size() != bytes) {
      out.data.resize(bytes);
    }
    uint8_t dummy[16];
    d2pc_cloud cloud;
    const int rc = d2pc_process_mono8_into(ctx_, msg->data.data(), msg->width, msg->height, msg->step,
                                           out.data.empty() ? dummy : out.data.data(), out.data.size(), &cloud);
    if (rc != D2PC_OK) {
      ROS_ERROR("d2pc_process_mono8_into: %s (%s)", d2pc_strerror(rc), d2pc_last_cuda_error(ctx_));
      return;
    }
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
    out.header.stamp = msg->header.stamp;
    out.header.frame_id = "/camera_optical_frame";
    pub_.publish(out);
  }

 private:
  ros::NodeHandle nh_;
  ros::Subscriber sub_;
  ros::Publisher pub_;
  d2pc_ctx *ctx_ = nullptr;
  sensor_msgs::PointCloud2 out_;   // persistent so that its data storage can stay page-locked
  uint8_t *registered_ = nullptr;
  int border_ = 40;
};

}  // namespace

int main(int argc, char **argv) {
  ros::init(argc, argv, "disparity_to_point_cloud");
  Disparity2PCloudGpu node;
  ros::spin();
  return 0;
}
