// euclidean_clusters.cpp -- cloud_geometry::nearest::extractEuclideanClusters on the B200.
// The reference calls the point_cloud_mapping function on the points above a detected table
// (cloud_tools/src/table_object_detector_passive.cpp:270-293: object_indices, object_cluster_tolerance_ 0.05,
// object_cluster_min_pts_ 30) and describes every resulting cluster with GRSD.  Here the `indices` subset is
// gathered, uploaded and clustered by cab_euclidean_clusters; labels come back in subset order and are mapped to
// the caller's point indices.
#include <algorithm>
#include <string>
#include <vector>

#include <cloud_algos/cloud_algos.h>
#include <point_cloud_mapping/geometry/nearest.h>

#include "cloud_algos_b200.h"

namespace
{
  std::string last_error;
  cloud_algos::GpuContext& gpu ()
  {
    static cloud_algos::GpuContext ctx;  // one context for the free function, created on first use
    return ctx;
  }
}

const std::string& cloud_geometry::nearest::lastEuclideanClusterError ()
{
  return last_error;
}

void cloud_geometry::nearest::extractEuclideanClusters (const sensor_msgs::PointCloud &points, const std::vector<int> &indices,
                                                        double tolerance, std::vector<std::vector<int> > &clusters, int nx_idx,
                                                        int ny_idx, int nz_idx, double eps_angle,
                                                        unsigned int min_pts_per_cluster)
{
  clusters.clear ();
  last_error.clear ();
  if (nx_idx != -1 || ny_idx != -1 || nz_idx != -1)
  {
    last_error = "extractEuclideanClusters: the normal-angle variant is not offloaded";
    ROS_ERROR ("%s", last_error.c_str ());
    return;
  }
  if (indices.empty ()) return;
  cab_ctx* ctx = gpu ().get (last_error);
  if (!ctx) { ROS_ERROR ("extractEuclideanClusters: %s", last_error.c_str ()); return; }

  const size_t n = indices.size ();
  std::vector<float> xyz (3 * n);
  for (size_t i = 0; i < n; ++i)
  {
    const geometry_msgs::Point32 &p = points.points.at (indices[i]);
    xyz[3 * i] = p.x; xyz[3 * i + 1] = p.y; xyz[3 * i + 2] = p.z;
  }
  std::vector<int32_t> labels (n, -1);
  int rc = cab_upload_cloud (ctx, &xyz[0], (int64_t) n, 3);
  int64_t nc = rc;
  if (rc == CAB_OK) nc = cab_euclidean_clusters (ctx, tolerance, (int32_t) min_pts_per_cluster, 0, &labels[0]);
  if (nc < 0)
  {
    last_error = cab_last_error (ctx);
    ROS_ERROR ("extractEuclideanClusters: %s", last_error.c_str ());
    return;
  }
  std::vector<int32_t> offsets ((size_t) nc + 1), members (n);
  cab_cluster_csr (&labels[0], (int64_t) n, (int32_t) nc, &offsets[0], &members[0]);
  clusters.resize ((size_t) nc);
  for (int64_t c = 0; c < nc; ++c)
  {
    std::vector<int> &r = clusters[c];
    r.reserve (offsets[c + 1] - offsets[c]);
    for (int32_t j = offsets[c]; j < offsets[c + 1]; ++j) r.push_back (indices[members[j]]);
    // the reference sorts every cluster's indices (and `indices` is ascending at its call sites anyway)
    std::sort (r.begin (), r.end ());
  }
}
