// cloud_tools/fit_sac_plane.h -- fitSACPlane of the table detectors
//   cloud_tools/src/table_object_detector_passive.cpp:621-659 (called per cluster of Z-parallel points at :415)
//   cloud_tools/src/table_object_detector_sr.cpp (same member)
// as a free function with the member's parameters plus the two members it reads (sac_distance_threshold_,
// clusters_min_pts_), backed by the B200 library (cab_fit_plane_msac).  The reference builds a
// sample_consensus::MSAC (setMaxIterations (500), setProbability (0.99)) over a SACModelPlane [point_cloud_mapping];
// here the hypotheses are scored on the GPU, 32 per launch.  There is no CPU fallback.
#ifndef CAB_CLOUD_TOOLS_FIT_SAC_PLANE_H
#define CAB_CLOUD_TOOLS_FIT_SAC_PLANE_H
#include <string>
#include <vector>

#include <sensor_msgs/PointCloud.h>

namespace cloud_tools
{
  /** Finds the best plane in the points `indices` of `points`: inliers (point indices within sac_distance_threshold of
    * the refined plane, in the order of `indices`) and coeff (a, b, c, d).  As in the reference the inliers are projected
    * onto the plane IN PLACE (model->projectPointsInPlace).  Returns -1 with empty outputs if there are fewer than
    * clusters_min_pts indices, 0 otherwise (empty outputs if the model has fewer than clusters_min_pts inliers).
    * `seed` starts the sample sequence (the reference draws with rand (); any sequence is as valid). */
  int fitSACPlane (sensor_msgs::PointCloud *points, std::vector<int> *indices, std::vector<int> &inliers,
                   std::vector<double> &coeff, double sac_distance_threshold, int clusters_min_pts, unsigned int seed = 1);
  /** Empty after a successful call, else the library's error message. */
  const std::string& lastFitSACPlaneError ();
}
#endif
