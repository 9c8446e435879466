// global_rsd.h -- cloud_algos::GlobalRSD.  The reference tree does not contain this plugin's
// header or source; its surface is recovered from the only caller,
// dyn_obj_store/src/table_memory_grsd.cpp:28,288,852,973-997: public fields min_voxel_pts_, step_,
// width_, publish_cloud_centroids_, publish_cloud_vrsd_, label_; rosparams width (0.03) and step (0)
// (dyn_obj_store/table_pipeline_grsd.launch:54-55); output = a one-point cloud whose channels
// f1..f21 hold the histogram, read at values.at(0).  The algorithm is the in-tree GRSD-21
// (color_chlac/include/color_chlac/grsd_colorCHLAC_tools.hpp:131-294) on the B200.
#ifndef CLOUD_ALGOS_GLOBAL_RSD_H
#define CLOUD_ALGOS_GLOBAL_RSD_H
#include <cloud_algos/cloud_algos.h>

namespace cloud_algos
{

class GlobalRSD : public CloudAlgo
{
 public:
  typedef sensor_msgs::PointCloud OutputType;
  typedef sensor_msgs::PointCloud InputType;

  // Options
  double width_;         // voxel (leaf) size; rosparam "width"
  int step_;             // 0: radius search around each voxel centroid (the only mode; rosparam "step")
  int min_voxel_pts_;    // voxels with fewer points are ignored; only 0 / 1 (= keep all) supported
  int label_;            // written to a "point_label" channel of the output if != -1
  bool publish_cloud_centroids_;  // keep the voxel centroids for getCentroids()
  bool publish_cloud_vrsd_;       // keep per-voxel r_min / r_max / surface type for getVRSD()
  double rsd_radius_min_;         // rsd_radius_search (grsd_colorCHLAC_tools.h:27)

  static std::string default_input_topic () {return std::string ("cloud_pcd");}
  static std::string default_output_topic () {return std::string ("cloud_grsd");}
  static std::string default_node_name () {return std::string ("global_rsd_node");}

  void init (ros::NodeHandle&);
  void pre ();
  void post ();
  std::vector<std::string> requires ();
  std::vector<std::string> provides ();
  std::string process (const boost::shared_ptr<const InputType>&);
  boost::shared_ptr<const OutputType> output ();
  /** Not in the reference (whose caller loops over the object clusters and calls process () once per cluster,
    * dyn_obj_store/src/table_memory_grsd.cpp:913-997): all clusters of a frame in ONE device call (cab_grsd_batch), which
    * is where the GPU's batched rate comes from.  outputs[i] is what process (clusters[i]) + output () would have
    * produced; returns "ok" or the first error. */
  std::string process_batch (const std::vector<boost::shared_ptr<const InputType> >& clusters,
                             std::vector<boost::shared_ptr<const OutputType> >& outputs);
  boost::shared_ptr<const sensor_msgs::PointCloud> getCentroids () {return cloud_centroids_;}
  boost::shared_ptr<const sensor_msgs::PointCloud> getVRSD () {return cloud_vrsd_;}

  GlobalRSD () : CloudAlgo (), width_ (0.03), step_ (0), min_voxel_pts_ (1), label_ (-1),
    publish_cloud_centroids_ (false), publish_cloud_vrsd_ (false), rsd_radius_min_ (0.01) {}

  ros::Publisher createPublisher (ros::NodeHandle& nh)
  {
    ros::Publisher p = nh.advertise<OutputType> (default_output_topic (), 5);
    return p;
  }
 private:
  ros::NodeHandle nh_;
  boost::shared_ptr<sensor_msgs::PointCloud> cloud_grsd_, cloud_centroids_, cloud_vrsd_;
  GpuContext gpu_;
};

}
#endif
