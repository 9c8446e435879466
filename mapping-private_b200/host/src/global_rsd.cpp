// global_rsd.cpp -- cloud_algos::GlobalRSD on the B200 (surface recovered from its caller, see the
// header; algorithm = extractGRSDSignature21, grsd_colorCHLAC_tools.hpp:131-294, hist_num == 1).
#include <cstdio>

#include <cloud_algos/cloud_algos.h>
#include <cloud_algos/global_rsd.h>

#include "cloud_algos_b200.h"

using namespace cloud_algos;

void GlobalRSD::init (ros::NodeHandle& nh)
{
  nh_ = nh;
}

void GlobalRSD::pre ()
{
  nh_.param ("width", width_, width_);
  nh_.param ("step", step_, step_);
  nh_.param ("min_voxel_pts", min_voxel_pts_, min_voxel_pts_);
  nh_.param ("label", label_, label_);
  nh_.param ("publish_cloud_centroids", publish_cloud_centroids_, publish_cloud_centroids_);
  nh_.param ("publish_cloud_vrsd", publish_cloud_vrsd_, publish_cloud_vrsd_);
}

void GlobalRSD::post ()
{
}

std::vector<std::string> GlobalRSD::requires ()
{
  std::vector<std::string> r;
  r.push_back ("x"); r.push_back ("y"); r.push_back ("z");
  r.push_back ("nx"); r.push_back ("ny"); r.push_back ("nz");
  return r;
}

std::vector<std::string> GlobalRSD::provides ()
{
  std::vector<std::string> p;
  for (int i = 1; i <= 21; ++i)
  {
    char name[8];
    std::snprintf (name, sizeof (name), "f%d", i);
    p.push_back (name);
  }
  return p;
}

std::string GlobalRSD::process (const boost::shared_ptr<const GlobalRSD::InputType>& cloud)
{
  output_valid_ = true;
  const int nxIdx = getChannelIndex (cloud, "nx");
  if (nxIdx == -1 || nxIdx + 2 >= (int) cloud->channels.size ())
  {
    ROS_ERROR ("[GlobalRSD] Provided point cloud does not have normals. Use the normal_estimation or mls_fit first!");
    output_valid_ = false;
    return std::string ("missing normals");
  }
  if (step_ != 0 || min_voxel_pts_ > 1)
  {
    output_valid_ = false;
    return std::string ("unsupported options: only step = 0 and min_voxel_pts <= 1 are implemented");
  }
  std::string err;
  cab_ctx* ctx = gpu_.get (err, true);
  if (!ctx) { output_valid_ = false; ROS_ERROR ("[GlobalRSD] %s", err.c_str ()); return err; }

  const size_t n = cloud->points.size ();
  const int32_t offsets[2] = {0, (int32_t) n};
  int32_t hist[21] = {0};
  const float vp[3] = {0, 0, 0};
  const float* xyz = n ? &cloud->points[0].x : 0;
  const float* nx = n ? &cloud->channels[nxIdx + 0].values[0] : 0;
  const float* ny = n ? &cloud->channels[nxIdx + 1].values[0] : 0;
  const float* nz = n ? &cloud->channels[nxIdx + 2].values[0] : 0;
  int rc = CAB_OK;
  if (n) rc = cab_grsd_batch (ctx, xyz, 3, offsets, 1, (float) width_, 0.f, rsd_radius_min_, 0, vp, nx, ny, nz, hist);
  if (rc != CAB_OK)
  {
    output_valid_ = false;
    err = std::string ("GRSD failed: ") + cab_last_error (ctx);
    ROS_ERROR ("[GlobalRSD] %s", err.c_str ());
    return err;
  }

  // one point (the cloud's centroid) carrying f1..f21 (read at values.at (0), table_memory_grsd.cpp:995-996)
  cloud_grsd_ = boost::shared_ptr<sensor_msgs::PointCloud> (new sensor_msgs::PointCloud ());
  cloud_grsd_->header = cloud->header;
  cloud_grsd_->points.resize (1);
  double cx = 0, cy = 0, cz = 0;
  for (size_t i = 0; i < n; ++i) { cx += cloud->points[i].x; cy += cloud->points[i].y; cz += cloud->points[i].z; }
  if (n) { cloud_grsd_->points[0].x = (float) (cx / n); cloud_grsd_->points[0].y = (float) (cy / n); cloud_grsd_->points[0].z = (float) (cz / n); }
  cloud_grsd_->channels.resize (21);
  for (int i = 0; i < 21; ++i)
  {
    char name[8];
    std::snprintf (name, sizeof (name), "f%d", i + 1);
    cloud_grsd_->channels[i].name = name;
    cloud_grsd_->channels[i].values.assign (1, (float) hist[i]);
  }
  if (label_ != -1)
  {
    cloud_grsd_->channels.resize (22);
    cloud_grsd_->channels[21].name = "point_label";
    cloud_grsd_->channels[21].values.assign (1, (float) label_);
  }

  cloud_centroids_.reset ();
  cloud_vrsd_.reset ();
  if ((publish_cloud_centroids_ || publish_cloud_vrsd_) && n)
  {
    int64_t voff[2] = {0, 0};
    const int64_t nv = cab_grsd_voxels (ctx, voff, 0, 0, 0, 0, 0);
    if (nv < 0)
    {
      output_valid_ = false;
      err = std::string ("GRSD voxel download failed: ") + cab_last_error (ctx);
      ROS_ERROR ("[GlobalRSD] %s", err.c_str ());
      return err;
    }
    std::vector<float> c (3 * (size_t) nv), rmin ((size_t) nv), rmax ((size_t) nv);
    std::vector<int32_t> lab ((size_t) nv);
    if (nv > 0) cab_grsd_voxels (ctx, voff, &c[0], &rmin[0], &rmax[0], &lab[0], nv);
    boost::shared_ptr<sensor_msgs::PointCloud> v (new sensor_msgs::PointCloud ());
    v->header = cloud->header;
    v->points.resize ((size_t) nv);
    for (int64_t i = 0; i < nv; ++i) { v->points[i].x = c[3 * i]; v->points[i].y = c[3 * i + 1]; v->points[i].z = c[3 * i + 2]; }
    if (publish_cloud_centroids_) cloud_centroids_ = boost::shared_ptr<sensor_msgs::PointCloud> (new sensor_msgs::PointCloud (*v));
    if (publish_cloud_vrsd_)
    {
      v->channels.resize (3);
      v->channels[0].name = "r_min"; v->channels[0].values.assign (rmin.begin (), rmin.end ());
      v->channels[1].name = "r_max"; v->channels[1].values.assign (rmax.begin (), rmax.end ());
      v->channels[2].name = "point_label"; v->channels[2].values.assign (lab.begin (), lab.end ());
      cloud_vrsd_ = v;
    }
  }
  return std::string ("ok");
}

boost::shared_ptr<const GlobalRSD::OutputType> GlobalRSD::output ()
  {return cloud_grsd_;}

std::string GlobalRSD::process_batch (const std::vector<boost::shared_ptr<const GlobalRSD::InputType> >& clusters,
                                      std::vector<boost::shared_ptr<const GlobalRSD::OutputType> >& outputs)
{
  outputs.clear ();
  output_valid_ = true;
  if (step_ != 0 || min_voxel_pts_ > 1)
  {
    output_valid_ = false;
    return std::string ("unsupported options: only step = 0 and min_voxel_pts <= 1 are implemented");
  }
  const size_t nc = clusters.size ();
  if (nc == 0) return std::string ("ok");
  std::vector<int32_t> offsets (nc + 1, 0);
  for (size_t c = 0; c < nc; ++c)
  {
    const int nxIdx = getChannelIndex (clusters[c], "nx");
    if (nxIdx == -1 || nxIdx + 2 >= (int) clusters[c]->channels.size ())
    {
      ROS_ERROR ("[GlobalRSD] Provided point cloud does not have normals. Use the normal_estimation or mls_fit first!");
      output_valid_ = false;
      return std::string ("missing normals");
    }
    offsets[c + 1] = offsets[c] + (int32_t) clusters[c]->points.size ();
  }
  const size_t n = (size_t) offsets[nc];
  std::vector<float> xyz (3 * n), nx (n), ny (n), nz (n);
  for (size_t c = 0; c < nc; ++c)
  {
    const InputType& cl = *clusters[c];
    const int nxIdx = getChannelIndex (clusters[c], "nx");
    for (size_t i = 0; i < cl.points.size (); ++i)
    {
      const size_t j = (size_t) offsets[c] + i;
      xyz[3 * j] = cl.points[i].x; xyz[3 * j + 1] = cl.points[i].y; xyz[3 * j + 2] = cl.points[i].z;
      nx[j] = cl.channels[nxIdx].values[i]; ny[j] = cl.channels[nxIdx + 1].values[i]; nz[j] = cl.channels[nxIdx + 2].values[i];
    }
  }
  std::string err;
  cab_ctx* ctx = gpu_.get (err, true);
  if (!ctx) { output_valid_ = false; ROS_ERROR ("[GlobalRSD] %s", err.c_str ()); return err; }
  std::vector<int32_t> hist (21 * nc, 0);
  const float vp[3] = {0, 0, 0};
  if (n)
  {
    const int rc = cab_grsd_batch (ctx, &xyz[0], 3, &offsets[0], (int32_t) nc, (float) width_, 0.f, rsd_radius_min_, 0, vp, &nx[0], &ny[0],
                                   &nz[0], &hist[0]);
    if (rc != CAB_OK)
    {
      output_valid_ = false;
      err = std::string ("GRSD failed: ") + cab_last_error (ctx);
      ROS_ERROR ("[GlobalRSD] %s", err.c_str ());
      return err;
    }
  }
  outputs.resize (nc);
  for (size_t c = 0; c < nc; ++c)
  {
    const InputType& cl = *clusters[c];
    boost::shared_ptr<sensor_msgs::PointCloud> o (new sensor_msgs::PointCloud ());
    o->header = cl.header;
    o->points.resize (1);
    double cx = 0, cy = 0, cz = 0;
    const size_t m = cl.points.size ();
    for (size_t i = 0; i < m; ++i) { cx += cl.points[i].x; cy += cl.points[i].y; cz += cl.points[i].z; }
    if (m) { o->points[0].x = (float) (cx / m); o->points[0].y = (float) (cy / m); o->points[0].z = (float) (cz / m); }
    o->channels.resize (label_ != -1 ? 22 : 21);
    for (int i = 0; i < 21; ++i)
    {
      char name[8];
      std::snprintf (name, sizeof (name), "f%d", i + 1);
      o->channels[i].name = name;
      o->channels[i].values.assign (1, (float) hist[21 * c + i]);
    }
    if (label_ != -1)
    {
      o->channels[21].name = "point_label";
      o->channels[21].values.assign (1, (float) label_);
    }
    outputs[c] = o;
  }
  return std::string ("ok");
}
