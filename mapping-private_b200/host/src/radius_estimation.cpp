// radius_estimation.cpp -- cloud_algos::LocalRadiusEstimation (RSD) on the B200.
// Same contract as cloud_algos/src/radius_estimation.cpp of the reference: rosparams read in
// pre() (:12-20), requires x,y,z,nx,ny,nz (:27-39), provides r_min,r_max,r_dif,point_label (:41-50),
// process() copies the input and appends those four channels (:75-95), returns "ok" or
// "missing normals" (:63-67,220).  Deliberate difference: a failing process() also clears
// output_valid_ (the reference leaves it set on "missing normals").
#include <cloud_algos/cloud_algos.h>
#include <cloud_algos/radius_estimation.h>

#include "cloud_algos_b200.h"

using namespace cloud_algos;

void LocalRadiusEstimation::init (ros::NodeHandle& nh)
{
  nh_ = nh;
}

void LocalRadiusEstimation::pre ()
{
  nh_.param ("radius", radius_, radius_);
  nh_.param ("max_nn", max_nn_, max_nn_);
  nh_.param ("plane_radius", plane_radius_, plane_radius_);
  nh_.param ("distance_div", distance_div_, distance_div_);
  nh_.param ("point_label", point_label_, point_label_);
  nh_.param ("rmin2curvature", rmin2curvature_, rmin2curvature_);
}

void LocalRadiusEstimation::post ()
{
}

std::vector<std::string> LocalRadiusEstimation::requires ()
{
  std::vector<std::string> r;
  r.push_back ("x"); r.push_back ("y"); r.push_back ("z");
  r.push_back ("nx"); r.push_back ("ny"); r.push_back ("nz");
  return r;
}

std::vector<std::string> LocalRadiusEstimation::provides ()
{
  std::vector<std::string> p;
  p.push_back ("r_min"); p.push_back ("r_max"); p.push_back ("r_dif"); p.push_back ("point_label");
  return p;
}

std::string LocalRadiusEstimation::process (const boost::shared_ptr<const LocalRadiusEstimation::InputType>& cloud)
{
  output_valid_ = true;
  const int nxIdx = getChannelIndex (cloud, "nx");
  if (nxIdx == -1 || nxIdx + 2 >= (int) cloud->channels.size ())
  {
    ROS_ERROR ("[LocalRadiusEstimation] Provided point cloud does not have normals. Use the normal_estimation or mls_fit first!");
    output_valid_ = false;
    return std::string ("missing normals");
  }
  std::string err;
  cab_ctx* ctx = gpu_.get (err);
  if (!ctx) { output_valid_ = false; ROS_ERROR ("[LocalRadiusEstimation] %s", err.c_str ()); return err; }

  ros::Time global_time = ros::Time::now ();
  cloud_radius_ = boost::shared_ptr<sensor_msgs::PointCloud> (new sensor_msgs::PointCloud (*cloud));
  const size_t n = cloud_radius_->points.size ();
  const int rIdx = (int) cloud_radius_->channels.size ();
  cloud_radius_->channels.resize (rIdx + 4);
  cloud_radius_->channels[rIdx + 0].name = "r_min";
  cloud_radius_->channels[rIdx + 1].name = "r_max";
  cloud_radius_->channels[rIdx + 2].name = "r_dif";
  const int labelIdx = rIdx + 3;
  cloud_radius_->channels[labelIdx].name = "point_label";
  for (int d = rIdx; d < rIdx + 4; ++d) cloud_radius_->channels[d].values.resize (n);

  const int cIdx = getChannelIndex (cloud, "curvature");
  if (rmin2curvature_ && cIdx == -1)
    ROS_ERROR ("[LocalRadiusEstimation] Overwriting of curvature values was requested but the channel doesn't exist!");

  // nx is followed by ny and nz (the reference's assumption, radius_estimation.cpp:68)
  const float* xyz = n ? &cloud->points[0].x : 0;
  const float* nx = n ? &cloud->channels[nxIdx + 0].values[0] : 0;
  const float* ny = n ? &cloud->channels[nxIdx + 1].values[0] : 0;
  const float* nz = n ? &cloud->channels[nxIdx + 2].values[0] : 0;
  float* rmin = n ? &cloud_radius_->channels[rIdx + 0].values[0] : 0;
  float* rmax = n ? &cloud_radius_->channels[rIdx + 1].values[0] : 0;
  int rc = cab_upload_cloud (ctx, xyz, (int64_t) n, 3);
  if (rc == CAB_OK) rc = cab_build_grid (ctx, (float) radius_);
  if (rc == CAB_OK) rc = cab_set_normals (ctx, nx, ny, nz);
  if (rc == CAB_OK) rc = cab_rsd (ctx, radius_, max_nn_, distance_div_, plane_radius_, 0, rmin, rmax);
  // r_dif is rounded from the double difference (radius_estimation.cpp:206), not from the rounded radii
  if (rc == CAB_OK && n) rc = cab_download_rdif (ctx, &cloud_radius_->channels[rIdx + 2].values[0]);
  if (rc != CAB_OK)
  {
    output_valid_ = false;
    err = std::string ("radius estimation failed: ") + cab_last_error (ctx);
    ROS_ERROR ("[LocalRadiusEstimation] %s", err.c_str ());
    return err;
  }
  for (size_t cp = 0; cp < n; ++cp)
  {
    if (rmin2curvature_ && cIdx != -1) cloud_radius_->channels[cIdx].values[cp] = rmin[cp];
    if (point_label_ != -1) cloud_radius_->channels[labelIdx].values[cp] = point_label_;
  }
  cab_timings tm;
  cab_profile (ctx, &tm);
  ROS_INFO ("[LocalRadiusEstimation] grid %g ms, radius estimation %g ms; processed point cloud in %g seconds.",
            tm.build_ms, tm.rsd_ms, (ros::Time::now () - global_time).toSec ());
  return std::string ("ok");
}

boost::shared_ptr<const LocalRadiusEstimation::OutputType> LocalRadiusEstimation::output ()
  {return cloud_radius_;}

#ifdef CREATE_NODE
// the <algo>_node executable of the reference's CMakeLists.txt:42-57 (cloud_algos.h:106-117)
int main (int argc, char* argv[])
{
  return cloud_algos::standalone_node <cloud_algos::LocalRadiusEstimation> (argc, argv);
}
#endif
