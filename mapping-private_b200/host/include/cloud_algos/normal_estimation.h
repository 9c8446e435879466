// normal_estimation.h -- cloud_algos::NormalEstimation, the plugin slot the reference advertises
// (cloud_algos/plugins.xml:3-7, include/cloud_algos/normal_estimation.h) but ships only as a stub
// (deprecated/normal_estimation.cpp returns an empty cloud).  Here it is real: radius-neighbourhood
// PCA normals with the semantics of pcl::NormalEstimation + setRadiusSearch as the reference calls it
// (color_chlac/include/color_chlac/grsd_colorCHLAC_tools.hpp:76-81), computed on the B200.
#ifndef CLOUD_ALGOS_NORMAL_ESTIMATION_H
#define CLOUD_ALGOS_NORMAL_ESTIMATION_H
#include <cloud_algos/cloud_algos.h>

namespace cloud_algos
{

class NormalEstimation : public CloudAlgo
{
 public:
  typedef sensor_msgs::PointCloud OutputType;
  typedef sensor_msgs::PointCloud InputType;

  // Options (rosparam keys = names without the trailing underscore)
  double radius_;   // neighbourhood radius (normals_radius_search = 0.02, grsd_colorCHLAC_tools.h:28)
  int max_nn_;      // keep only the max_nn nearest neighbours; <= 0: all
  double vp_x_, vp_y_, vp_z_;  // viewpoint the normals are flipped towards

  static std::string default_input_topic () {return std::string ("cloud_pcd");}
  static std::string default_output_topic () {return std::string ("cloud_normals");}
  static std::string default_node_name () {return std::string ("normal_estimation_node");}

  void init (ros::NodeHandle&);
  void pre ();
  void post ();
  std::vector<std::string> requires ();
  std::vector<std::string> provides ();
  std::string process (const boost::shared_ptr<const InputType>&);
  boost::shared_ptr<const OutputType> output ();

  NormalEstimation () : CloudAlgo (), radius_ (0.02), max_nn_ (0), vp_x_ (0), vp_y_ (0), vp_z_ (0) {}

  ros::Publisher createPublisher (ros::NodeHandle& nh)
  {
    ros::Publisher p = nh.advertise<OutputType> (default_output_topic (), 5);
    return p;
  }
 private:
  ros::NodeHandle nh_;
  boost::shared_ptr<sensor_msgs::PointCloud> cloud_normals_;
  GpuContext gpu_;
};

}
#endif
