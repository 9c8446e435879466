// radius_estimation.h -- cloud_algos::LocalRadiusEstimation (RSD) with the reference's public
// surface (cloud_algos/include/cloud_algos/radius_estimation.h:26-112): same option fields and
// defaults, same topics, same call sequence.  The kd-tree and the materialised neighbour lists
// (radius_estimation.h:108-110) are gone; the work is done by libcloudalgos_b200 on the B200.
#ifndef CLOUD_ALGOS_RADIUS_H
#define CLOUD_ALGOS_RADIUS_H
#include <cloud_algos/cloud_algos.h>

namespace cloud_algos
{

class LocalRadiusEstimation : public CloudAlgo
{
 public:
  typedef sensor_msgs::PointCloud OutputType;
  typedef sensor_msgs::PointCloud InputType;

  // Options (defaults: radius_estimation.h:81-86)
  double radius_;       // search radius for getting the nearest neighbors
  int max_nn_;          // maximum number of nearest neighbors to consider
  double plane_radius_; // radius value to set for planes
  int distance_div_;    // number of divisions for distance discretization
  int point_label_;     // class label to write, or -1
  bool rmin2curvature_; // overwrite curvature values with r_min if enabled

  static std::string default_input_topic () {return std::string ("cloud_pcd");}
  static std::string default_output_topic () {return std::string ("cloud_radius");}
  static std::string default_node_name () {return std::string ("radius_estimation_node");}

  void init (ros::NodeHandle&);
  void pre ();
  void post ();
  std::vector<std::string> requires ();
  std::vector<std::string> provides ();
  std::string process (const boost::shared_ptr<const InputType>&);
  boost::shared_ptr<const OutputType> output ();

  // kept for source compatibility: the reference drops its kd-tree here (radius_estimation.h:64-73)
  void clear () {}

  LocalRadiusEstimation () : CloudAlgo (),
    radius_ (0.03), max_nn_ (150), plane_radius_ (0.1), distance_div_ (10), point_label_ (-1), rmin2curvature_ (false) {}

  ros::Publisher createPublisher (ros::NodeHandle& nh)
  {
    ros::Publisher p = nh.advertise<OutputType> (default_output_topic (), 5);
    return p;
  }
 private:
  ros::NodeHandle nh_;
  boost::shared_ptr<sensor_msgs::PointCloud> cloud_radius_;
  GpuContext gpu_;
};

}
#endif
