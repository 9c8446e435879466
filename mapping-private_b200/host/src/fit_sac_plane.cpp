// fit_sac_plane.cpp -- fitSACPlane (cloud_tools/src/table_object_detector_passive.cpp:621-659) on the B200.
#include <cstdint>
#include <string>
#include <vector>

#include <cloud_algos/cloud_algos.h>
#include <cloud_tools/fit_sac_plane.h>

#include "cloud_algos_b200.h"

namespace
{
  std::string last_error;
  cloud_algos::GpuContext& gpu ()
  {
    static cloud_algos::GpuContext ctx;
    return ctx;
  }
}

const std::string& cloud_tools::lastFitSACPlaneError ()
{
  return last_error;
}

int cloud_tools::fitSACPlane (sensor_msgs::PointCloud *points, std::vector<int> *indices, std::vector<int> &inliers,
                              std::vector<double> &coeff, double sac_distance_threshold, int clusters_min_pts, unsigned int seed)
{
  last_error.clear ();
  inliers.resize (0);
  coeff.resize (0);
  if ((int) indices->size () < clusters_min_pts) return (-1);  // :625-630
  cab_ctx* ctx = gpu ().get (last_error);
  if (!ctx) { ROS_ERROR ("fitSACPlane: %s", last_error.c_str ()); return (0); }

  const size_t n = points->points.size (), m = indices->size ();
  std::vector<float> xyz (3 * n);
  for (size_t i = 0; i < n; ++i)
  {
    xyz[3 * i] = points->points[i].x; xyz[3 * i + 1] = points->points[i].y; xyz[3 * i + 2] = points->points[i].z;
  }
  std::vector<int32_t> idx (indices->begin (), indices->end ());
  // the sample sequence: max_iterations + 1 triples of distinct positions, a 64-bit LCG (Knuth's MMIX constants)
  const int max_iterations = 500;  // :636
  std::vector<int32_t> triples (3 * (size_t) (max_iterations + 1));
  uint64_t state = 0x9E3779B97F4A7C15ull ^ (uint64_t) seed;
  for (size_t t = 0; t < triples.size (); t += 3)
    for (int k = 0; k < 3; ++k)
    {
      for (;;)
      {
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        const int32_t pos = (int32_t) ((state >> 33) % m);
        if ((k > 0 && pos == triples[t]) || (k > 1 && pos == triples[t + 1])) { if (m < 3) break; else continue; }
        triples[t + k] = pos;
        break;
      }
    }
  double c[4];
  std::vector<int32_t> in (m);
  std::vector<float> proj (3 * m);
  int64_t nin = cab_fit_plane_msac (ctx, &xyz[0], (int64_t) n, 3, &idx[0], (int64_t) m, sac_distance_threshold, max_iterations, 0.99,
                                    &triples[0], (int64_t) (triples.size () / 3), c, &in[0], &proj[0], (int64_t) m, 0, 0);
  if (nin < 0)
  {
    last_error = cab_last_error (ctx);
    ROS_ERROR ("fitSACPlane: %s", last_error.c_str ());
    return (0);
  }
  if ((int) nin < clusters_min_pts) return (0);  // :643-649
  coeff.assign (c, c + 4);
  inliers.assign (in.begin (), in.begin () + nin);
  for (int64_t i = 0; i < nin; ++i)  // :657 projectPointsInPlace
  {
    geometry_msgs::Point32 &p = points->points[inliers[i]];
    p.x = proj[3 * i]; p.y = proj[3 * i + 1]; p.z = proj[3 * i + 2];
  }
  return (0);
}
