// noise_removal.h -- cloud_algos::StatisticalNoiseRemoval on the B200.
// Same public surface as cloud_algos/include/cloud_algos/noise_removal.h of the reference (:23-92):
// options alpha_ (3), neighborhood_size_ (10, counts the point itself), min_nr_pts_ (0); topics
// cloud_pcd -> cloud_denoise, node statistical_noise_removal_node.  The kd-tree members of the
// reference (kdtree_, points_indices_, points_sqr_distances_) are replaced by the GPU context.
#ifndef CLOUD_ALGOS_NOISE_REMOVAL_H
#define CLOUD_ALGOS_NOISE_REMOVAL_H
#include <cloud_algos/cloud_algos.h>

namespace cloud_algos
{

class StatisticalNoiseRemoval : public CloudAlgo
{
 public:
  typedef sensor_msgs::PointCloud OutputType;
  typedef sensor_msgs::PointCloud InputType;

  // Options
  double alpha_;          // discard points with average nearest neighbors distance further than alpha_*STD
  int neighborhood_size_; // number of nearest neighbors (including self) to consider
  int min_nr_pts_;        // minimum number of points in the cloud that is still acceptable

  static std::string default_input_topic () {return std::string ("cloud_pcd");}
  static std::string default_output_topic () {return std::string ("cloud_denoise");}
  static std::string default_node_name () {return std::string ("statistical_noise_removal_node");}

  void init (ros::NodeHandle&);
  void pre ();
  void post ();
  std::vector<std::string> requires ();
  std::vector<std::string> provides ();
  std::string process (const boost::shared_ptr<const InputType>&);
  boost::shared_ptr<const OutputType> output ();

  // nothing is carried from one process() to the next (the reference frees its kd-tree here, :59-68)
  void clear () {}

  StatisticalNoiseRemoval () : CloudAlgo (), alpha_ (3), neighborhood_size_ (10), min_nr_pts_ (0) {}

  ros::Publisher createPublisher (ros::NodeHandle& nh)
  {
    ros::Publisher p = nh.advertise<OutputType> (default_output_topic (), 5);
    return p;
  }
 private:
  ros::NodeHandle nh_;
  boost::shared_ptr<sensor_msgs::PointCloud> cloud_denoise_;
  GpuContext gpu_;
};

}
#endif
