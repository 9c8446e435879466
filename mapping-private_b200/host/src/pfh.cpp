// pfh.cpp -- cloud_algos::PointFeatureHistogram on the B200.
// Contract of cloud_algos/src/pfh.cpp of the reference: rosparams read in pre() (:12-24), requires x, y, z, nx,
// ny, nz (:31-43), provides f1..f<nr_bins> (+ point_label) with nr_bins = quantum * (3 or 4) (:45-76); process()
// returns "missing normals" without the nx channel (:81-95), copies the cloud, appends the feature channels
// (:127-163) and fills them (:205-350), the combined n-D histogram (combine_, :251-265) included.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include <cloud_algos/cloud_algos.h>
#include <cloud_algos/pfh.h>

#include "cloud_algos_b200.h"

using namespace cloud_algos;

void PointFeatureHistogram::init (ros::NodeHandle& nh)
{
  nh_ = nh;
}

void PointFeatureHistogram::pre ()
{
  nh_.param ("radius", radius_, radius_);
  nh_.param ("max_nn", max_nn_, max_nn_);
  nh_.param ("quantum", quantum_, quantum_);
  nh_.param ("use_dist", use_dist_, use_dist_);
  nh_.param ("combine", combine_, combine_);
  nh_.param ("differential", differential_, differential_);
  nh_.param ("check_flip", check_flip_, check_flip_);
  nh_.param ("abs_angles", abs_angles_, abs_angles_);
  nh_.param ("average", average_, average_);
  nh_.param ("point_label", point_label_, point_label_);
}

void PointFeatureHistogram::post ()
{
}

std::vector<std::string> PointFeatureHistogram::requires ()
{
  std::vector<std::string> requires;
  requires.push_back ("x"); requires.push_back ("y"); requires.push_back ("z");
  requires.push_back ("nx"); requires.push_back ("ny"); requires.push_back ("nz");
  return requires;
}

std::vector<std::string> PointFeatureHistogram::provides ()
{
  nr_features_ = use_dist_ ? 4 : 3;
  if (combine_) nr_bins_ = (int) ceil (pow (quantum_, nr_features_));
  else nr_bins_ = quantum_ * nr_features_;
  std::vector<std::string> provides;
  for (int i = 0; i < nr_bins_; i++)
  {
    char dim_name[16];
    std::snprintf (dim_name, sizeof (dim_name), "f%d", i + 1);
    provides.push_back (dim_name);
  }
  if (point_label_ != -1) provides.push_back ("point_label");
  return provides;
}

std::string PointFeatureHistogram::process (const boost::shared_ptr<const PointFeatureHistogram::InputType>& cloud)
{
  clear ();
  output_valid_ = true;
  const int nxIdx = getChannelIndex (cloud, "nx");
  if (nxIdx == -1 || nxIdx + 2 >= (int) cloud->channels.size ())
  {
    ROS_ERROR ("[PointFeatureHistogram] Provided point cloud does not have normals. Use the normal_estimation or mls_fit first!");
    output_valid_ = false;
    return std::string ("missing normals");
  }
  nr_features_ = use_dist_ ? 4 : 3;
  nr_bins_ = combine_ ? (int) ceil (pow (quantum_, nr_features_)) : quantum_ * nr_features_;  // :98-108
  std::string err;
  cab_ctx* ctx = gpu_.get (err);
  if (!ctx) { output_valid_ = false; ROS_ERROR ("[PointFeatureHistogram] %s", err.c_str ()); return err; }

  ros::Time global_time = ros::Time::now ();
  cloud_pfh_ = boost::shared_ptr<sensor_msgs::PointCloud> (new sensor_msgs::PointCloud ());
  cloud_pfh_->header   = cloud->header;
  cloud_pfh_->points   = cloud->points;
  cloud_pfh_->channels = cloud->channels;
  const int fIdx = (int) cloud_pfh_->channels.size ();
  cloud_pfh_->channels.resize (fIdx + nr_bins_ + (point_label_ != -1 ? 1 : 0));
  const size_t n = cloud_pfh_->points.size ();
  for (int i = 0; i < nr_bins_; i++)
  {
    char dim_name[16];
    std::snprintf (dim_name, sizeof (dim_name), "f%d", i + 1);
    cloud_pfh_->channels[fIdx + i].name = dim_name;
    cloud_pfh_->channels[fIdx + i].values.resize (n, 0.0);
  }
  if (point_label_ != -1)
  {
    cloud_pfh_->channels[fIdx + nr_bins_].name = "point_label";
    cloud_pfh_->channels[fIdx + nr_bins_].values.assign (n, (float) point_label_);
  }

  const int flags = (use_dist_ ? CAB_PFH_USE_DIST : 0) | (differential_ ? CAB_PFH_DIFFERENTIAL : 0) |
                    (check_flip_ ? CAB_PFH_CHECK_FLIP : 0) | (abs_angles_ ? CAB_PFH_ABS_ANGLES : 0) |
                    (average_ ? CAB_PFH_AVERAGE : 0) | (combine_ ? CAB_PFH_COMBINE : 0);
  std::vector<float> hist (n * (size_t) nr_bins_);
  int rc = CAB_OK;
  if (n)
  {
    // nx is followed by ny and nz (the reference's assumption, pfh.cpp:96)
    rc = cab_upload_cloud (ctx, &cloud->points[0].x, (int64_t) n, 3);
    if (rc == CAB_OK) rc = cab_build_grid (ctx, (float) radius_);
    if (rc == CAB_OK) rc = cab_set_normals (ctx, &cloud->channels[nxIdx].values[0], &cloud->channels[nxIdx + 1].values[0],
                                            &cloud->channels[nxIdx + 2].values[0]);
    if (rc == CAB_OK) rc = cab_pfh (ctx, radius_, max_nn_, quantum_, flags, &hist[0]);
  }
  if (rc != CAB_OK)
  {
    output_valid_ = false;
    err = std::string ("PFH failed: ") + cab_last_error (ctx);
    ROS_ERROR ("[PointFeatureHistogram] %s", err.c_str ());
    return err;
  }
  for (int b = 0; b < nr_bins_; b++)
  {
    std::vector<float>& v = cloud_pfh_->channels[fIdx + b].values;
    for (size_t cp = 0; cp < n; cp++) v[cp] = hist[cp * nr_bins_ + b];
  }
  ROS_INFO ("[PointFeatureHistogram] PFH done in %g seconds.", (ros::Time::now () - global_time).toSec ());
  return std::string ("ok");
}

PointFeatureHistogram::OutputType PointFeatureHistogram::output ()
  {return *(cloud_pfh_.get ());}

#ifdef CREATE_NODE
// the <algo>_node executable of the reference's CMakeLists.txt:42-57 (cloud_algos.h:106-117)
int main (int argc, char* argv[])
{
  return cloud_algos::standalone_node <cloud_algos::PointFeatureHistogram> (argc, argv);
}
#endif
