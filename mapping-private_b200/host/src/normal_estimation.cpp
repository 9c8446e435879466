// normal_estimation.cpp -- cloud_algos::NormalEstimation on the B200 (see the header).
#include <cloud_algos/cloud_algos.h>
#include <cloud_algos/normal_estimation.h>

#include "cloud_algos_b200.h"

using namespace cloud_algos;

void NormalEstimation::init (ros::NodeHandle& nh)
{
  nh_ = nh;
}

void NormalEstimation::pre ()
{
  nh_.param ("radius", radius_, radius_);
  nh_.param ("max_nn", max_nn_, max_nn_);
  nh_.param ("vp_x", vp_x_, vp_x_);
  nh_.param ("vp_y", vp_y_, vp_y_);
  nh_.param ("vp_z", vp_z_, vp_z_);
}

void NormalEstimation::post ()
{
}

std::vector<std::string> NormalEstimation::requires ()
{
  std::vector<std::string> r;
  r.push_back ("x"); r.push_back ("y"); r.push_back ("z");
  return r;
}

std::vector<std::string> NormalEstimation::provides ()
{
  std::vector<std::string> p;
  p.push_back ("nx"); p.push_back ("ny"); p.push_back ("nz"); p.push_back ("curvature");
  return p;
}

std::string NormalEstimation::process (const boost::shared_ptr<const NormalEstimation::InputType>& cloud)
{
  output_valid_ = true;
  std::string err;
  cab_ctx* ctx = gpu_.get (err);
  if (!ctx) { output_valid_ = false; ROS_ERROR ("[NormalEstimation] %s", err.c_str ()); return err; }

  // the input is never modified: copy it and add / reuse the normal channels
  cloud_normals_ = boost::shared_ptr<sensor_msgs::PointCloud> (new sensor_msgs::PointCloud (*cloud));
  const size_t n = cloud_normals_->points.size ();
  const char* names[4] = {"nx", "ny", "nz", "curvature"};
  int idx[4];
  for (int c = 0; c < 4; ++c)
  {
    idx[c] = getChannelIndex (*cloud_normals_, names[c]);
    if (idx[c] == -1)
    {
      idx[c] = (int) cloud_normals_->channels.size ();
      cloud_normals_->channels.resize (idx[c] + 1);
      cloud_normals_->channels[idx[c]].name = names[c];
    }
    cloud_normals_->channels[idx[c]].values.resize (n);
  }

  const float vp[3] = {(float) vp_x_, (float) vp_y_, (float) vp_z_};
  std::vector<float> n4 (4 * n);
  const float* xyz = n ? &cloud->points[0].x : 0;
  int rc = cab_upload_cloud (ctx, xyz, (int64_t) n, 3);
  if (rc == CAB_OK) rc = cab_build_grid (ctx, (float) radius_);
  if (rc == CAB_OK) rc = cab_normals (ctx, (float) radius_, max_nn_, vp, n ? &n4[0] : 0);
  if (rc != CAB_OK)
  {
    output_valid_ = false;
    err = std::string ("normal estimation failed: ") + cab_last_error (ctx);
    ROS_ERROR ("[NormalEstimation] %s", err.c_str ());
    return err;
  }
  for (size_t i = 0; i < n; ++i)
    for (int c = 0; c < 4; ++c)
      cloud_normals_->channels[idx[c]].values[i] = n4[4 * i + c];
  return std::string ("ok");
}

boost::shared_ptr<const NormalEstimation::OutputType> NormalEstimation::output ()
  {return cloud_normals_;}

#ifdef CREATE_NODE
// the <algo>_node executable of the reference's CMakeLists.txt:42-57 (cloud_algos.h:106-117)
int main (int argc, char* argv[])
{
  return cloud_algos::standalone_node <cloud_algos::NormalEstimation> (argc, argv);
}
#endif
