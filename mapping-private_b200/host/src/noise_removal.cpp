// noise_removal.cpp -- cloud_algos::StatisticalNoiseRemoval on the B200.
// Contract of cloud_algos/src/noise_removal.cpp of the reference: rosparams alpha, neighborhood_size,
// min_nr_pts read in pre() (:12-17); requires and provides x, y, z (:24-43); process() rejects
// neighborhood_size < 2 or alpha < 0 (:51-56) and clouds smaller than the neighbourhood (:57-62),
// keeps the points whose mean neighbour distance lies within alpha standard deviations of the cloud's
// mean (:84-146), copies their channel values and fails the output size check against min_nr_pts
// (:152-158).  The k-NN search and the per-point means run on the GPU (cab_statistical_outliers).
#include <vector>

#include <cloud_algos/cloud_algos.h>
#include <cloud_algos/noise_removal.h>

#include "cloud_algos_b200.h"

using namespace cloud_algos;

void StatisticalNoiseRemoval::init (ros::NodeHandle& nh)
{
  nh_ = nh;
}

void StatisticalNoiseRemoval::pre ()
{
  nh_.param ("alpha", alpha_, alpha_);
  nh_.param ("neighborhood_size", neighborhood_size_, neighborhood_size_);
  nh_.param ("min_nr_pts", min_nr_pts_, min_nr_pts_);
}

void StatisticalNoiseRemoval::post ()
{
}

std::vector<std::string> StatisticalNoiseRemoval::requires ()
{
  std::vector<std::string> requires;
  requires.push_back ("x");
  requires.push_back ("y");
  requires.push_back ("z");
  return requires;
}

std::vector<std::string> StatisticalNoiseRemoval::provides ()
{
  std::vector<std::string> provides;
  provides.push_back ("x");
  provides.push_back ("y");
  provides.push_back ("z");
  return provides;
}

std::string StatisticalNoiseRemoval::process (const boost::shared_ptr<const StatisticalNoiseRemoval::InputType>& cloud)
{
  clear ();
  output_valid_ = true;
  if (neighborhood_size_ < 2 || alpha_ < 0)
  {
    if (verbosity_level_ > -2) ROS_ERROR ("[StatisticalNoiseRemoval] A STD limit of %g and/or a neighborhood of size %d makes no sense!", alpha_, neighborhood_size_);
    output_valid_ = false;
    return std::string ("ERROR: Not enough neighbors requested!");
  }
  if (neighborhood_size_ > (int) cloud->points.size ())
  {
    if (verbosity_level_ > -2) ROS_ERROR ("[StatisticalNoiseRemoval] %d nearest neighbors (including self) requested, but only %d points in total!", neighborhood_size_, (int) cloud->points.size ());
    output_valid_ = false;
    return std::string ("ERROR: Not enough points in the cloud (or too many neighbors requested)!");
  }
  std::string err;
  cab_ctx* ctx = gpu_.get (err);
  if (!ctx) { output_valid_ = false; ROS_ERROR ("[StatisticalNoiseRemoval] %s", err.c_str ()); return err; }

  ros::Time global_time = ros::Time::now ();
  const size_t n = cloud->points.size ();
  std::vector<uint8_t> keep (n, 0);
  double mean = 0, stddev = 0;
  int rc = cab_upload_cloud (ctx, &cloud->points[0].x, (int64_t) n, 3);
  int64_t kept = 0;
  if (rc == CAB_OK)
  {
    kept = cab_statistical_outliers (ctx, neighborhood_size_, alpha_, 0.f, &keep[0], 0, &mean, &stddev);
    if (kept < 0) rc = (int) kept;
  }
  if (rc != CAB_OK)
  {
    output_valid_ = false;
    err = std::string ("noise removal failed: ") + cab_last_error (ctx);
    ROS_ERROR ("[StatisticalNoiseRemoval] %s", err.c_str ());
    return err;
  }
  if (verbosity_level_ > 0) ROS_INFO ("[StatisticalNoiseRemoval] Computed mean (%g) and std (%g).", mean, stddev);

  // Copy the necessary data from the original PCD (:124-143)
  cloud_denoise_ = boost::shared_ptr<sensor_msgs::PointCloud> (new sensor_msgs::PointCloud ());
  cloud_denoise_->header = cloud->header;
  cloud_denoise_->points.reserve ((size_t) kept);
  cloud_denoise_->channels = cloud->channels;
  int point_count = 0;
  for (size_t cp = 0; cp < n; cp++)
  {
    if (keep[cp])
    {
      cloud_denoise_->points.push_back (cloud->points[cp]);
      for (unsigned d = 0; d < cloud->channels.size (); d++)
        cloud_denoise_->channels[d].values[point_count] = cloud->channels[d].values[cp];
      point_count++;
    }
  }
  for (unsigned d = 0; d < cloud->channels.size (); d++)
    cloud_denoise_->channels[d].values.resize (point_count);
  if (verbosity_level_ > 0) ROS_INFO ("[StatisticalNoiseRemoval] Selected %d/%d points (%g%%); cleared cloud in %g seconds.", point_count, (int) n, 100 * (point_count / (double) n), (ros::Time::now () - global_time).toSec ());

  if (point_count < min_nr_pts_)
  {
    if (verbosity_level_ > -1) ROS_WARN ("[StatisticalNoiseRemoval] Number of points smaller than the threshold of %d. Invalid output!", min_nr_pts_);
    output_valid_ = false;
    return std::string ("output size check failed (see min_nr_pts parameter)");
  }
  output_valid_ = true;
  return std::string ("ok");
}

boost::shared_ptr<const StatisticalNoiseRemoval::OutputType> StatisticalNoiseRemoval::output ()
  {return cloud_denoise_;}

#ifdef CREATE_NODE
// the <algo>_node executable of the reference's CMakeLists.txt:42-57 (cloud_algos.h:106-117)
int main (int argc, char* argv[])
{
  return cloud_algos::standalone_node <cloud_algos::StatisticalNoiseRemoval> (argc, argv);
}
#endif
