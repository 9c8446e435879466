/*
 * oracle.h -- CPU restatement of the reference's normals -> RSD -> GRSD path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (mapping-private_b200/,
 * include/) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker or
 * as the reported CPU baseline.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - RSD arithmetic, get_type, the 26-offset table, transition counting and the
 *     21-bin packing are restated from in-tree reference sources (cited per
 *     function); no reference test pins their outputs -> "parity unpinned" for
 *     normals and radii.
 *   - voxel occupancy + neighbour lookup are pinned by known answers decoded from
 *     color_chlac/demos/shape_data/noiseless_*_GRSD_CCHLAC.pcd (tests/golden/).
 *   - radius search, PCA normals and VoxelGrid live in third-party code that is
 *     absent from /root/reference (point_cloud_mapping/ANN, PCL 1.5.1 pinned at
 *     hough_segmentation/CMakeLists.txt:5, FLANN); their published semantics are
 *     restated and the call sites are cited.
 *
 * All entry points are plain C so tests can bind them with ctypes.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* RSD behaviour flags (DESIGN.md "RSD variants"). Default 0 = in-tree
 * LocalRadiusEstimation (cloud_algos/src/radius_estimation.cpp:140-215). */
#define ORC_RSD_REF_IS_NEAREST 1 /* reference element = nearest surface point (PCL RSDEstimation) */
#define ORC_RSD_SEED_BIN0 2      /* bin 0 starts at angle 0 (newer PCL) */
#define ORC_RSD_SCALE_SORT 4     /* r_min*=1.1, r_max*=0.9, then order (newer PCL) */

/* Squared distance under the documented epsilon rule: fp32, (dx*dx+dy*dy)+dz*dz,
 * no FMA contraction. */
float orc_d2(const float* a, const float* b);

/* Radius search over `surface` for each query position.
 *   neighbour  <=>  orc_d2(surface[j], query) <= (float)r*(float)r
 * Results per query sorted by (d2, index) ascending and truncated to max_nn
 * (max_nn <= 0 = unlimited) -- restates kdtree_->radiusSearch(cp, radius_, idx,
 * d2, max_nn_) as used at radius_estimation.cpp:120 [EXTERNAL semantics].
 * offsets has nq+1 entries. idx/d2 may be NULL (count only). Returns the total
 * number of neighbours; if it exceeds cap only offsets are filled. */
int64_t orc_radius_search(const float* surface_xyz, int n, const float* query_xyz, int nq,
                          double r, int max_nn, int64_t* offsets, int32_t* idx, float* d2,
                          int64_t cap, int nthreads);

/* O(N*Q) brute-force variant of the above, for cross-checks on small inputs. */
int64_t orc_radius_search_brute(const float* surface_xyz, int n, const float* query_xyz, int nq,
                                double r, int max_nn, int64_t* offsets, int32_t* idx, float* d2,
                                int64_t cap);

/* PCA normals, pcl::NormalEstimation semantics [EXTERNAL] as called at
 * grsd_colorCHLAC_tools.hpp:76-81: neighbourhood within r (self included),
 * smallest-eigenvalue eigenvector of the covariance, curvature = l0/(l0+l1+l2),
 * flipped so that n.(vp - p) >= 0; fewer than 3 neighbours -> NaN.
 * out_n4: n x {nx,ny,nz,curvature}; out_k (optional): neighbour counts. */
int orc_normals(const float* xyz, int n, double r, int max_nn, const float* vp,
                float* out_n4, int32_t* out_k, int nthreads);
/* Same normals plus the eigenvalue gap (l1 - l0) / trace of every point (NaN where the normal is NaN). */
int orc_normals_gap(const float* xyz, int n, double r, int max_nn, const float* vp, float* out_n4, float* out_gap,
                    int nthreads);

/* RSD for every point of a cloud with per-point normals (stride floats apart).
 * Restates radius_estimation.cpp:140-215.  Outputs are fp32 like the reference's
 * channels (r_dif = (float)(double r_max - double r_min)). */
int orc_rsd(const float* xyz, const float* normals, int normal_stride, int n, double r,
            int max_nn, int ndiv, double plane_radius, int flags, float* r_min, float* r_max,
            float* r_dif, int nthreads);

/* RSD at arbitrary query positions over a surface cloud (the
 * pcl::RSDEstimation + setSearchSurface use at grsd_colorCHLAC_tools.hpp:164-180). */
int orc_rsd_queries(const float* surface_xyz, const float* normals, int normal_stride, int n,
                    const float* query_xyz, int nq, double r, int max_nn, int ndiv,
                    double plane_radius, int flags, float* r_min, float* r_max, int nthreads);

/* Same as orc_rsd but organised like the reference: build a kd-tree, materialise
 * every neighbour list, then run the estimation loop; the three phases are timed
 * where the reference logs them (radius_estimation.cpp:108,125,216).
 * phase_s[3] = {tree build, neighbour search, estimation} seconds. 1 thread. */
int orc_rsd_ref_faithful(const float* xyz, const float* normals, int normal_stride, int n,
                         double r, int max_nn, int ndiv, double plane_radius, float* r_min,
                         float* r_max, double* phase_s);

/* pcl::VoxelGrid with setSaveLeafLayout(true) [EXTERNAL], as used at
 * grsd_colorCHLAC_tools.hpp:94-100.  Call with centroids == NULL to get sizes.
 *   min_b[3], div_b[3]; returns number of occupied voxels V.
 *   centroids: V x 3 (mean xyz = fp32 sum in cloud order times 1.0f / count, the Eigen 3.0-3.2 `/=` of the
 *              reference's era, pinned by its shipped feature vectors; voxels ordered by linear index x-fastest)
 *   layout: div_b[0]*div_b[1]*div_b[2] ints, voxel -> centroid index or -1
 *   counts: V point counts (optional) */
int orc_voxel_grid(const float* xyz, int n, float leaf, int32_t* min_b, int32_t* div_b,
                   float* centroids, int32_t* layout, int32_t* counts);

/* grsd_colorCHLAC_tools.hpp:104-116 */
int orc_get_type(float min_radius, float max_radius);

/* 26-neighbour offsets in the reference's order (grsd_colorCHLAC_tools.hpp:187-222);
 * out: 26 x 3 ints. */
void orc_offsets26(int32_t* out);

/* Transition counting + packing for precomputed voxel labels
 * (grsd_colorCHLAC_tools.hpp:230-292 with hist_num == 1).
 *   transition36: 6x6 row-major M(src, nbr); hist21: upper triangle. */
int orc_grsd_transitions(const float* centroids, int nvox, const int32_t* types, float leaf,
                         const int32_t* min_b, const int32_t* div_b, const int32_t* layout,
                         int32_t* transition36, int32_t* hist21);

/* Whole GRSD-21 recipe for one cluster (exampleRSD.cpp:50-93 +
 * extractGRSDSignature21): normals(r_normals) -> voxel grid(leaf) -> RSD at the
 * centroids with r = max(rsd_radius_min, leaf/2*sqrt(3)), ndiv = 5, plane_radius
 * = 0.2, reference element = nearest surface point -> labels -> transitions.
 * normals_in may be NULL (then computed with viewpoint vp).  Optional debug
 * outputs (may be NULL): labels (cap_vox), radii (cap_vox x 2), nvox. */
int orc_grsd21(const float* xyz, const float* normals_in, int normal_stride, int n, float leaf,
               double r_normals, double rsd_radius_min, int rsd_flags, const float* vp,
               int32_t* hist21, int32_t* labels, float* radii, int32_t cap_vox,
               int32_t* nvox_out, int nthreads);

/* Subdivision variant (hist_num > 1), grsd_colorCHLAC_tools.hpp:140-161,233-246.
 * Call with hist21 == NULL to get subdiv_b[3]; returns hist_num (0 if offsets
 * exceed the grid, -1 on invalid subdivision size). */
int orc_grsd21_subdiv(const float* centroids, int nvox, const int32_t* types, float leaf,
                      const int32_t* min_b, const int32_t* div_b, const int32_t* layout,
                      int subdivision_size, int off_x, int off_y, int off_z, int32_t* subdiv_b,
                      int32_t* hist21);

/* Mean normal of every voxel.  pcl::VoxelGrid with downsample_all_data averages every field of
 * the point type, the normals included, and does not re-normalise them (comment at
 * grsd_colorCHLAC_tools.hpp:558) [EXTERNAL].  out: V x 3 = fp32 sum in (voxel, input index) order times
 * 1.0f / count (the arithmetic of orc_voxel_grid's centroids); voxel order as orc_voxel_grid.  Returns V. */
int orc_voxel_normals(const float* xyz, const float* normals, int normal_stride, int n, float leaf,
                      float* out);

/* The three transition signatures of grsd_colorCHLAC_tools.hpp for precomputed voxel labels, with
 * the subdivision / sliding-box mode (subdivision_size > 0, voxel offsets off_*):
 *   ORC_SIG_GRSD21      extractGRSDSignature21      (:131-294)  21 bins per histogram
 *   ORC_SIG_GRSD325     extractGRSDSignature325     (:305-451)  325 bins: src + 5*nbr + 25*offset_id
 *                       over the 13 half offsets, EMPTY neighbours ignored (:427-428)
 *   ORC_SIG_PLUSGRSD110 extractPlusGRSDSignature110 (:462-668)  7 normal-angle bins x 15 + 5
 *                       to-empty; needs cent_normals (V x 3, un-normalised voxel means), which are
 *                       normalised in fp32 like Eigen's normalize() (:559-560)
 * Call with hist == NULL to get subdiv_b[3] and the number of histograms; returns hist_num
 * (0 if the offsets exceed the grid, -1 on an invalid subdivision size, -2 on a bad kind).
 * hist: hist_num x dim int32 (the reference emits them as floats, optionally x NORMALIZE_GRSD). */
#define ORC_SIG_GRSD21 0
#define ORC_SIG_GRSD325 1
#define ORC_SIG_PLUSGRSD110 2
int orc_grsd_signature(int kind, const float* centroids, const float* cent_normals, int nvox,
                       const int32_t* types, float leaf, const int32_t* min_b, const int32_t* div_b,
                       const int32_t* layout, int subdivision_size, int off_x, int off_y, int off_z,
                       int32_t* subdiv_b, int32_t* hist);

/* libsvm C-SVC / RBF prediction with svm-scale style feature scaling, as
 * cloud_algos::SVMClassification::process does per point (svm_classification.cpp:134-155):
 *   value_i = scaleFeature(i, (double)feature_i, ranges, lower, upper)   (svm_classification.h:68-86;
 *             skipped when fmin == NULL)
 *   class   = svm_predict(model, nodes)        [EXTERNAL: libsvm, manifest.xml:28, not vendored]
 * svm_predict for C_SVC restated from the published algorithm: k(x, sv) = exp(-gamma * sum (x-sv)^2)
 * accumulated in index order, one-vs-one decision values sum_i coef*k - rho in libsvm's pair order,
 * vote, first maximum wins, label[argmax].  All in double.
 * features: n x dim floats; labels / nr_sv: nr_class; rho: nr_class*(nr_class-1)/2;
 * sv_coef: (nr_class-1) x total_sv; sv: total_sv x dim dense.  out: n predicted labels (as float,
 * like the point_class channel); dec (optional): n x nr_class*(nr_class-1)/2 decision values. */
int orc_svm_predict(const float* features, int64_t n, int dim, int nr_class, int total_sv, double gamma,
                    const int32_t* labels, const int32_t* nr_sv, const double* rho, const double* sv_coef,
                    const double* sv, double lower, double upper, const double* fmin, const double* fmax,
                    float* out, double* dec);

/* cloud_algos::StatisticalNoiseRemoval (cloud_algos/src/noise_removal.cpp:84-136):
 *   avg[cp] = sum_{ni=1}^{k-1} (double)sqrtf(d2[cp][ni]) / (k-1) over the k nearest neighbours of cp
 *             (cp itself is the first and is skipped, :104-111), d2 under the documented rule, ties
 *             by input index [kdtree_->nearestKSearch: EXTERNAL semantics];
 *   mean / stddev of avg over the cloud (:113-121); keep cp iff |avg - mean| < alpha * stddev (:131).
 * Non-finite points get avg = NaN, are left out of the statistics and are never kept (the reference
 * has no such points).  orc_knn_mean_distance returns -2 when the cloud has fewer than k finite
 * points (:57-62), -1 for k < 2 (:51-56). */
int orc_knn_mean_distance(const float* xyz, int n, int k, double* avg, int nthreads);
int64_t orc_noise_filter(const double* avg, int n, double alpha, uint8_t* keep, double* mean_out,
                         double* stddev_out);

/* cloud_algos::PointFeatureHistogram (cloud_algos/src/pfh.cpp:78-366, pair features at
 * cloud_algos/include/cloud_algos/pfh.h:102-238): "star" pair features between every point and its
 * neighbours within `radius` (<= max_nn nearest, self first and skipped), one 1-D histogram of `quantum`
 * bins per feature (alpha, beta, gamma [, delta]) with increments of 100 / k, then -- with
 * ORC_PFH_AVERAGE, the FPFH step -- the 1/d2-weighted average of the neighbours' histograms (:303-333) and,
 * with ORC_PFH_DIFFERENTIAL, bin-to-bin differences (:337-350).  ORC_PFH_COMBINE: the combined n-D histogram mode
 * (combine_ = true).  out: n x nr_bins floats, point-major, nr_bins = quantum * (3 or 4), or quantum ^ (3 or 4).
 * Quirks kept: a point whose only neighbour is itself gets 0/0 = NaN with ORC_PFH_AVERAGE; a duplicate
 * neighbour (d2 == 0) yields an invalid pair (increment spread over all bins, :277-287) and an infinite
 * weight. */
#define ORC_PFH_USE_DIST 1
#define ORC_PFH_DIFFERENTIAL 2
#define ORC_PFH_CHECK_FLIP 4
#define ORC_PFH_ABS_ANGLES 8
#define ORC_PFH_AVERAGE 16
#define ORC_PFH_COMBINE 32 /* combine_: one n-D histogram of quantum^features bins (pfh.cpp:47-57, 239-258, 279-281) */
int orc_pfh(const float* xyz, const float* normals, int normal_stride, int n, double radius, int max_nn,
            int quantum, int flags, float* out, int nthreads);
/* One pair (pfh.h:102-238): returns 0 if the pair is invalid, else 1 and alpha, beta, gamma, delta in f[4]
 * (already normalised to [0, 1]).  d2 is the cached squared distance (:246). */
int orc_pfh_pair(const float* ps, const float* ns, const float* pt, const float* nt, float d2, double max_dist,
                 int check_flip, int abs_angles, double* f);

/* ---- colour half of VOSCH: rotation-invariant Color-CHLAC / C3-HLAC, 117 bins ---------------------------
 * color_chlac/include/color_chlac/color_chlac.hpp:1471-1528 (computeColorCHLAC), :1565-1743 (the four add functions),
 * :1745-1782 (normalisation), setColor :148-166, setVoxelFilter :181-210; called by extractC3HLACSignature117 and
 * extractVOSCH (grsd_colorCHLAC_tools.hpp:787-843; VOSCH = the 20 GRSD bins followed by these 117).
 * c3 = 1: C3HLAC_RI_Estimation (sin / cos colour coding), 0: ColorCHLAC_RI_Estimation (v, 255 - v).
 * rgb: one packed 0x00RRGGBB per voxel (orc_voxel_colors).  Call with hist == NULL for subdiv_b / hist_num; returns
 * hist_num (0: offsets exceed the grid, -1: invalid subdivision size, -2: negative colour threshold).
 * hist: hist_num x 117 floats, accumulated and normalised exactly in the reference's order. */
int orc_color_chlac117(int c3, const float* centroids, const uint32_t* rgb, int nvox, float leaf, const int32_t* min_b,
                       const int32_t* div_b, const int32_t* layout, int thR, int thG, int thB, int subdivision_size,
                       int off_x, int off_y, int off_z, int32_t* subdiv_b, float* hist);
/* pcl::VoxelGrid's voxel colour [EXTERNAL]: r, g, b summed and divided as floats, truncated; voxel order as
 * orc_voxel_grid.  Returns V. */
int orc_voxel_colors(const float* xyz, const uint32_t* rgb, int n, float leaf, uint32_t* out_rgb);

/* k-NN normals (table_object_detector_passive.cpp:668-714, cylinder_fit_algo.cpp:138-203): orc_normals' PCA over the
 * k nearest points (query included, ties by index).  Returns -1 for k < 3, -2 for fewer than k finite points. */
int orc_normals_knn(const float* xyz, int n, int k, const float* vp, float* out_n4, int nthreads);

/* ---- Euclidean clustering (the step that produces the segmented object clusters GRSD runs on) --------
 * cloud_geometry::nearest::extractEuclideanClusters(points, indices, tolerance, clusters, -1, -1, -1, -1, min_pts)
 * as called at cloud_tools/src/table_object_detector_passive.cpp:293,567 and table_object_detector_sr.cpp:370.
 * The function itself lives in point_cloud_mapping [EXTERNAL, absent]; its published algorithm is restated:
 * for i ascending over the unprocessed points: breadth-first growth from i through radiusSearch(tolerance)
 * (neighbour <=> d2 <= tolerance^2, this repo's epsilon rule), every reached point marked processed; the
 * region is kept iff it has >= min_pts points (and <= max_pts if max_pts > 0); kept regions are emitted in the
 * order of their seeds (= their smallest index) with their indices sorted ascending.
 * labels[n]: cluster id in that order, -1 for dropped regions and non-finite points.  Returns the cluster count. */
int orc_euclidean_clusters(const float* xyz, int n, double tolerance, int min_pts, int max_pts, int32_t* labels);
/* fitSACPlane (cloud_tools/src/table_object_detector_passive.cpp:621-659): MSAC plane through a given sample sequence,
 * refined by least squares, inliers within the threshold and their projections.  [parity unpinned: point_cloud_mapping's
 * sample_consensus is not in the tree and draws its samples with rand()] */
int64_t orc_fit_plane_msac(const float* xyz, int64_t n, const int32_t* indices, int64_t n_idx, double threshold,
                           int32_t max_iterations, double probability, const int32_t* triples, int64_t n_triples,
                           double coeff[4], int32_t* inliers, float* projected_xyz, int32_t* iterations_run,
                           int32_t* best_iteration);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
