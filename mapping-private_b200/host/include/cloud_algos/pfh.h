// pfh.h -- cloud_algos::PointFeatureHistogram on the B200.
// Same public surface as cloud_algos/include/cloud_algos/pfh.h of the reference (:23-100): options radius_
// (0.03), max_nn_ (100), quantum_ (9), use_dist_, combine_, differential_, check_flip_ (true), abs_angles_,
// average_ (true), point_label_ (-1) -- the defaults produce FPFHs; topics cloud_pcd -> cloud_pfh, node
// pfh_node; output() returns the cloud by value like the reference's (:62).  The pair features
// (getPointPairFeatures, :102-238) and the histogram loops run on the GPU through cab_pfh.
#ifndef CLOUD_ALGOS_PFH_H
#define CLOUD_ALGOS_PFH_H
#include <cloud_algos/cloud_algos.h>

#include <algorithm>
#include <cmath>

namespace cloud_algos
{

class PointFeatureHistogram : public CloudAlgo
{
 public:
  typedef sensor_msgs::PointCloud OutputType;
  typedef sensor_msgs::PointCloud InputType;

  // Options
  double radius_;     // search radius for getting the nearest neighbors
  int max_nn_;        // maximum number of nearest neighbors to consider
  int quantum_;       // number of divisions in a feature's definition interval
  bool use_dist_;     // enable to use distance as a feature
  bool combine_;      // enable to count co-occurrences of features (not implemented on the GPU path)
  bool differential_; // enable to let histogram values for a feature be relative to the previous one - if not combined
  bool check_flip_;   // enable to make source-target selection consistent
  bool abs_angles_;   // enable to use absolute values of angles instead of the 'directional' values
  bool average_;      // enable to create the final histogram by a weighted average of the neighboring ones
  int point_label_;   // set the value for the class label (-1 if not known / to be classified)

  static std::string default_input_topic () {return std::string ("cloud_pcd");}
  static std::string default_output_topic () {return std::string ("cloud_pfh");}
  static std::string default_node_name () {return std::string ("pfh_node");}

  void init (ros::NodeHandle&);
  void pre ();
  void post ();
  std::vector<std::string> requires ();
  std::vector<std::string> provides ();
  std::string process (const boost::shared_ptr<const InputType>&);
  OutputType output ();

  void clear () {}

  PointFeatureHistogram () : CloudAlgo (), radius_ (0.03), max_nn_ (100), quantum_ (9), use_dist_ (false), combine_ (false),
    differential_ (false), check_flip_ (true), abs_angles_ (false), average_ (true), point_label_ (-1), nr_features_ (3), nr_bins_ (27) {}

  // Compute the index of each feature - merge 1 into the last [s,s+1) interval (pfh.h:96-99)
  static inline int getFeatureIndice (int &quantum, double &feature)
  {
    return std::max (0, std::min (quantum - 1, (int) std::floor (quantum * feature)));
  }

  ros::Publisher createPublisher (ros::NodeHandle& nh)
  {
    ros::Publisher p = nh.advertise<OutputType> (default_output_topic (), 5);
    return p;
  }
 private:
  ros::NodeHandle nh_;
  boost::shared_ptr<sensor_msgs::PointCloud> cloud_pfh_;
  GpuContext gpu_;
  int nr_features_;
  int nr_bins_;
};

}
#endif
