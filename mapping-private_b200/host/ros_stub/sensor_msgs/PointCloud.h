#pragma once
// Stand-in for sensor_msgs/PointCloud.h (ROS 1): points[] AoS + named float channels (SoA).
#include <string>
#include <vector>
#include <boost/shared_ptr.hpp>
namespace std_msgs {
struct Header {
  unsigned int seq = 0;
  double stamp = 0;
  std::string frame_id;
};
}  // namespace std_msgs
namespace geometry_msgs {
struct Point32 {
  float x = 0, y = 0, z = 0;
};
}  // namespace geometry_msgs
namespace sensor_msgs {
struct ChannelFloat32 {
  std::string name;
  std::vector<float> values;
};
struct PointCloud {
  std_msgs::Header header;
  std::vector<geometry_msgs::Point32> points;
  std::vector<ChannelFloat32> channels;
};
typedef boost::shared_ptr<PointCloud> PointCloudPtr;
typedef boost::shared_ptr<const PointCloud> PointCloudConstPtr;
}  // namespace sensor_msgs
