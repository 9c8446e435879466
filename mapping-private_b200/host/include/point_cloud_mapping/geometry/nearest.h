// point_cloud_mapping/geometry/nearest.h -- the one function of cloud_geometry::nearest that sits on the way
// into the GRSD path, with the signature its callers use
//   cloud_tools/src/table_object_detector_passive.cpp:293,567
//   cloud_tools/src/table_object_detector_sr.cpp:370
// backed by the B200 library (cab_euclidean_clusters).  Put this directory in front of the include path of
// the real point_cloud_mapping package to reroute those calls; everything else of that package stays as it is.
#ifndef CAB_POINT_CLOUD_MAPPING_GEOMETRY_NEAREST_H
#define CAB_POINT_CLOUD_MAPPING_GEOMETRY_NEAREST_H
#include <vector>

#include <sensor_msgs/PointCloud.h>

namespace cloud_geometry
{
namespace nearest
{
  /** Decomposes the points `indices` of `points` into Euclidean clusters: two points belong to one cluster iff a
    * chain of points with consecutive distances <= tolerance joins them.  Clusters come in the order of their
    * smallest position in `indices`, each as ascending point indices; clusters with fewer than min_pts_per_cluster
    * points are dropped.  nx_idx / ny_idx / nz_idx / eps_angle (the normal-angle test of the region-growing
    * callers) must be -1: that variant is not offloaded and the call then returns no cluster and sets
    * lastEuclideanClusterError ().  Runs on the GPU; there is no CPU fallback. */
  void extractEuclideanClusters (const sensor_msgs::PointCloud &points, const std::vector<int> &indices, double tolerance,
                                 std::vector<std::vector<int> > &clusters, int nx_idx, int ny_idx, int nz_idx,
                                 double eps_angle, unsigned int min_pts_per_cluster = 1);
  /** Empty after a successful call, else the library's error message. */
  const std::string& lastEuclideanClusterError ();
}
}
#endif
