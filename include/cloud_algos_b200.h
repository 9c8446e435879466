/*
 * cloud_algos_b200.h -- C ABI of libcloudalgos_b200.so, the B200 (sm_100a) implementation of
 * the normals -> RSD -> GRSD hot path of cloud_algos.
 *
 * This is the drop-in boundary: plain C types, caller-allocated outputs, no STL / torch / ROS
 * types.  The C++ plugin shim (mapping-private_b200/host), bench.py and the tests bind exactly
 * these symbols.  Each entry point cites the reference interface it replaces (paths relative to
 * the reference tree).  There is no CPU fallback: every call fails with CAB_ERR_CUDA when no
 * sm_100a device is usable.
 *
 * Threading: a cab_ctx owns one CUDA stream (plus a copy stream used by cab_normals_rsd) and a grow-only device arena; a context must not
 * be used from two threads at once, different contexts are independent (the reference runs
 * each plugin instance on the single ros::spin() thread, cloud_algos.h:106-117).
 * Every call is synchronous: results are valid when it returns.
 */
#ifndef CLOUD_ALGOS_B200_H
#define CLOUD_ALGOS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAB_OK 0
#define CAB_ERR_ARG (-1)   /* bad argument / call order */
#define CAB_ERR_CUDA (-2)  /* CUDA runtime error, no device, wrong architecture */
#define CAB_ERR_OOM (-3)   /* device allocation failed / grid too large */
#define CAB_ERR_STATE (-4) /* required stage has not been run (no cloud, no grid, no normals) */

/* RSD behaviour flags.  0 = in-tree LocalRadiusEstimation
 * (cloud_algos/src/radius_estimation.cpp:140-215). */
#define CAB_RSD_SEED_BIN0 2  /* bin 0 starts at angle 0 (newer pcl::computeRSD) */
#define CAB_RSD_SCALE_SORT 4 /* r_min*=1.1, r_max*=0.9, then ordered (newer pcl::computeRSD) */
#define CAB_STEP_INPUT_ORDER 0x100 /* cab_step_normals_rsd on a context outside a group: the RSD kernel also leaves normals +
                                      curvature (float4) and r_min, r_max (float2) of every point in INPUT order on the device
                                      (cab_device_ptr CAB_BUF_NRM_INPUT_RANGE / CAB_BUF_RSD_INPUT_RANGE: the channels
                                      radius_estimation.cpp:204-214 appends) -- what a group's step leaves rank by rank */

typedef struct cab_ctx cab_ctx;

typedef struct cab_config {
  int32_t device;       /* CUDA device ordinal */
  int32_t exact;        /* 1: fp64 neighbour sums + fp64 acos (bit-reproducible on lattice inputs); 0: fp32 fast path */
  int64_t max_table_cells; /* budget for the dense cell table (0 = default 2^28) */
} cab_config;

typedef struct cab_timings {
  float build_ms;   /* last cab_build_grid: bounds + keys + radix sort + reorder + tables */
  float normals_ms; /* last cab_normals kernel(s) */
  float rsd_ms;     /* last cab_rsd kernel(s) */
  float grsd_ms;    /* last cab_grsd_batch, all kernels */
  float h2d_ms, d2h_ms; /* copies inside the last upload / download calls */
  int64_t n_points, n_valid, n_packets, n_rows, n_cells;
  int64_t neighbour_sum; /* sum over queries of in-radius neighbours of the last normals/rsd pass */
  int64_t candidate_sum; /* sum over queries of candidates tested in that pass */
  int64_t kernel_launches; /* kernels launched by this library since cab_create (own + CUB) */
  int64_t n_sorted; /* points radix-sorted by the last cab_build_grid: n_valid, or only the rows the shard, its halo
                       and their candidates lie in when the context is one shard of several (cab_set_shard) */
  float knn_ms;       /* last cab_knn_mean_distance / cab_statistical_outliers: all grid rounds */
  int32_t knn_rounds; /* grids built by it (the cell edge doubles until every query has its k neighbours) */
  float pfh_ms;       /* last cab_pfh: pair-feature, averaging and finishing kernels */
  float cluster_ms;   /* last cab_euclidean_clusters: union, statistics and labelling kernels (without the grid build) */
  float exchange_ms;  /* last cab_step_normals_rsd of a group: end of the RSD kernel -> every rank's results have arrived */
  float step_ms;      /* last cab_step_normals_rsd: first kernel of the build -> end of the step, on the device */
  int32_t shard_mode; /* last cab_build_grid: 0 whole cloud; 1 slab, halo normals recomputed; 2 slab, halo normals exchanged
                         with the neighbouring ranks (cab_step_normals_rsd of a group whose slabs are two layers thick) */
} cab_timings;

/* ---- lifetime ------------------------------------------------------------------------ */
int cab_create(const cab_config* cfg, cab_ctx** out);
void cab_destroy(cab_ctx* ctx);
/* Message of the last failing call on ctx (ctx may be NULL for cab_create failures). */
const char* cab_last_error(const cab_ctx* ctx);

/* ---- cloud ---------------------------------------------------------------------------
 * Replaces the deep copy + `new cloud_kdtree::KdTreeANN(*cloud)` input stage
 * (radius_estimation.cpp:75-78,103-109): points[] as AoS floats, `stride` floats apart
 * (3 for geometry_msgs::Point32, 4 for PCL PointXYZ).  Non-finite points are kept at their
 * index but never become neighbours. */
int cab_upload_cloud(cab_ctx* ctx, const float* xyz, int64_t n, int32_t stride);
/* Same, for a batch of independent segmented clusters: offsets[nclusters+1] into xyz.
 * Neighbourhoods never cross cluster boundaries (table_memory_grsd.cpp:913 loops clusters). */
int cab_upload_clusters(cab_ctx* ctx, const float* xyz, int64_t n, int32_t stride,
                        const int32_t* offsets, int32_t nclusters);
/* Device-resident variant (xyz already in HBM, e.g. a torch tensor's data_ptr). */
int cab_set_cloud_device(cab_ctx* ctx, const float* d_xyz, int64_t n, int32_t stride);

/* Builds the search structure that replaces the kd-tree (radius_estimation.cpp:107;
 * pcl::KdTreeFLANN at grsd_colorCHLAC_tools.hpp:79,175): device radix sort of
 * (row, x) keys, cell offsets, 32-query packets.  `cell` must be >= every radius used later.
 * The cell table is dense: a cloud too spread out for it at this cell size (cfg.max_table_cells) is
 * indexed with coarser cells -- same results, more candidates tested per query. */
int cab_build_grid(cab_ctx* ctx, float cell);

/* Query sharding for multi-GPU runs (the reference is single-threaded, radius_estimation.cpp:139): this
 * context answers the queries of one slab of the cloud -- whole rows of the grid, cut so that every rank's cost (its
 * own rows in both passes plus the normals of one layer of halo rows around them) is the same; the cuts follow from a
 * fixed sample of the cloud and are identical on every rank.  Default (0,1).  With world > 1 the next cab_build_grid
 * keys, sorts and tabulates only the rows this rank reads (own + halo + the halo's candidates): the device arrays
 * (CAB_BUF_*) are then local to the slab, cab_download returns this shard's rows only -- use cab_download_sorted /
 * CAB_OUT_SHARD_SORTED, or the cab_comm_* calls below, which concatenate the ranks' results.  Every shard's results
 * equal the unsharded ones bit for bit.  One cloud only: a batch of clusters (cab_upload_clusters) shards by cluster. */
int cab_set_shard(cab_ctx* ctx, int32_t rank, int32_t world);
/* Element range [begin, end) of this context's own queries in its (local) sorted arrays. */
int cab_shard_range(cab_ctx* ctx, int64_t* begin, int64_t* end);

/* ---- normals -------------------------------------------------------------------------
 * Replaces pcl::NormalEstimation::compute with setRadiusSearch(r)
 * (grsd_colorCHLAC_tools.hpp:76-81; the cloud_algos/NormalEstimation plugin slot,
 * plugins.xml:3-7).  out: n x {nx, ny, nz, curvature} in input order, may be NULL to keep the
 * result on the device only.  max_nn <= 0: unlimited. */
int cab_normals(cab_ctx* ctx, float r, int32_t max_nn, const float vp[3], float* nxyz_curv);
/* Provide normals computed elsewhere (the LocalRadiusEstimation plugin receives them as the
 * nx,ny,nz channels, radius_estimation.cpp:58-68).  Three SoA arrays of n floats. */
int cab_set_normals(cab_ctx* ctx, const float* nx, const float* ny, const float* nz);

/* ---- RSD -----------------------------------------------------------------------------
 * Replaces HOT LOOP 1 + 2 of LocalRadiusEstimation::process (radius_estimation.cpp:118-202):
 * radius search (<= max_nn nearest, self excluded), per-distance-bin min/max normal angle,
 * the two least-squares fits capped at plane_radius.  r_min / r_max: n floats each in input
 * order, may be NULL. */
int cab_rsd(cab_ctx* ctx, double r, int32_t max_nn, int32_t ndiv, double plane_radius,
            int32_t flags, float* r_min, float* r_max);

/* ---- normals + RSD in one call --------------------------------------------------------
 * The NormalEstimation -> LocalRadiusEstimation chain of sample_pipeline.yaml
 * (cloud_algos/sample_pipeline.yaml; exampleRSD.cpp:50-93 runs the same two stages back to back)
 * without the host round trip between the plugins: the normals are copied to the host on a second
 * stream while the RSD kernel runs.  Same results as cab_normals((float)r) followed by cab_rsd(r).
 * layout CAB_OUT_INPUT_ORDER: nxyz_curv n x 4, out_a = r_min[n], out_b = r_max[n] in input order
 * (input_index unused).  layout CAB_OUT_SHARD_SORTED (multi-GPU): this context's cab_shard_range
 * only, in sorted order: nxyz_curv m x 4, out_a = m x {r_min, r_max}, input_index[m] (out_b unused).
 * Any output pointer may be NULL.  Host buffers should be page-locked for the overlap to happen.
 * (Input-order layout: the two pass kernels store every query's result at its input index themselves, next to the
 * sorted-order arrays the second pass reads, so no permutation pass stands between a kernel and its copy.) */
#define CAB_OUT_INPUT_ORDER 0
#define CAB_OUT_SHARD_SORTED 1
int cab_normals_rsd(cab_ctx* ctx, double r, int32_t max_nn_normals, const float vp[3], int32_t max_nn_rsd,
                    int32_t ndiv, double plane_radius, int32_t flags, int32_t layout, float* nxyz_curv,
                    float* out_a, float* out_b, int32_t* input_index);

/* ---- parity / debug ------------------------------------------------------------------
 * Neighbour index sets of queries [q0, q1) (input order) as the radius search returns them
 * (radius_estimation.cpp:120), unsorted.  offsets: q1-q0+1 entries.  idx/d2 may be NULL
 * (count only); returns total or <0. */
int64_t cab_neighbors_debug(cab_ctx* ctx, float r, int32_t max_nn, int64_t q0, int64_t q1,
                            int64_t* offsets, int32_t* idx, float* d2, int64_t cap);

/* ---- GRSD ----------------------------------------------------------------------------
 * Replaces getVoxelGrid + extractGRSDSignature21 (grsd_colorCHLAC_tools.hpp:94-100,131-294)
 * for a batch of clusters, the work of cloud_algos/GlobalRSD as called from
 * table_memory_grsd.cpp:974-996.  normals: n x 3 floats (SoA arrays) or all NULL to compute
 * them with radius r_normals.  RSD radius = max(rsd_radius_min, leaf/2*sqrt(3)), nr_subdiv 5,
 * plane_radius 0.2, reference element = nearest surface point.
 * hist21: nclusters x 21 int32 (upper triangle of the 6x6 transition matrix). */
int cab_grsd_batch(cab_ctx* ctx, const float* xyz, int32_t stride, const int32_t* offsets,
                   int32_t nclusters, float leaf, float r_normals, double rsd_radius_min,
                   int32_t rsd_flags, const float vp[3], const float* nx, const float* ny,
                   const float* nz, int32_t* hist21);
/* Per-voxel results of the last cab_grsd_batch (voxels ordered by cluster, then linear voxel
 * index).  vox_offsets: nclusters+1. Any pointer may be NULL. Returns total voxels.
 * Centroids follow pcl::VoxelGrid of the reference's era: fp32 sums in cloud order times 1.0f / count (Eigen 3.0-3.2's
 * `/=`), and getNeighborCentroidIndices looks a centroid's voxel up by floor(c / leaf) -- both pinned by the reference's
 * shipped color_chlac/demos/shape_data vectors. */
int64_t cab_grsd_voxels(cab_ctx* ctx, int64_t* vox_offsets, float* centroids_xyz, float* r_min,
                        float* r_max, int32_t* labels, int64_t cap);

/* The other transition signatures of grsd_colorCHLAC_tools.hpp, computed from the voxels, labels and
 * leaf layouts the last cab_grsd_batch left on the device:
 *   CAB_SIG_GRSD21       extractGRSDSignature21      (:131-294)  21 bins
 *   CAB_SIG_GRSD325      extractGRSDSignature325     (:305-451)  325 bins, rotation variant
 *   CAB_SIG_PLUSGRSD110  extractPlusGRSDSignature110 (:462-668)  110 bins, voxel-normal angle
 * subdivision_size > 0 is the sliding-box mode (:140-161, :233-246; used by
 * color_voxel_recognition_2/.../search_new.h:62-76): one histogram per box of subdivision_size^3
 * voxels, boxes starting at voxel offset (off_x, off_y, off_z) of every cluster's grid; a cluster
 * whose grid is smaller than the offsets yields no histogram (the reference returns Zero, :150-153).
 * hist_offsets: nclusters + 1 (first histogram of every cluster); subdiv_b: nclusters x 3 boxes per
 * axis; hist: total x dim int32 counts (the reference emits them as floats, optionally times
 * NORMALIZE_GRSD).  Any pointer may be NULL; returns the total number of histograms or < 0. */
#define CAB_SIG_GRSD21 0
#define CAB_SIG_GRSD325 1
#define CAB_SIG_PLUSGRSD110 2
int64_t cab_grsd_signatures(cab_ctx* ctx, int32_t kind, int32_t subdivision_size, int32_t off_x, int32_t off_y,
                            int32_t off_z, int64_t* hist_offsets, int32_t* subdiv_b, int32_t* hist, int64_t cap);

/* GRSD of ONE large cloud -- the cloud of cab_upload_cloud / cab_set_cloud_device / cab_comm_upload_cloud --, the
 * getVoxelGrid + computeGRSD + extractGRSDSignature21 chain of grsd_colorCHLAC_tools.hpp:64-294 on a single cluster, and
 * its multi-GPU form: on a sharded context (cab_set_shard / a group) every rank voxelises the whole cloud
 * (cheap, replicated) but computes normals for its slab of rows and the RSD radii + surface-type labels of the voxels
 * whose centroid lies in its own rows; the ranks' labels are merged (every voxel has exactly one owner: the sum of
 * label + 1 | 0), every rank counts the neighbour transitions of its own voxels, and one small int32 all-reduce sums the
 * 6 x 6 transition matrices (their integer sums are order independent: bit-exact against one GPU).
 *   cab_grsd_cloud             all of it in a group (cab_comm_init*: the two merges are cab_comm_allreduce_i32) or on an
 *                              unsharded context; hist21 = the whole cloud's GRSD-21 on every rank.  Afterwards
 *                              cab_grsd_signatures (GRSD-21 with subdivisions, GRSD-325) works the same way: own voxels,
 *                              then the all-reduce of the hist_num x dim counts.  PlusGRSD is not available on a
 *                              sharded cloud (it needs the normals of every point of a voxel).
 *   cab_grsd_cloud_labels      step 1 alone: returns the number of voxels; labels_plus1[v] = label + 1 for this rank's
 *   cab_grsd_cloud_set_labels  voxels, 0 otherwise.  The application merges (sums) the ranks' arrays by its own means,
 *                              gives the merged labels back, and sums the partial histograms of cab_grsd_signatures. */
int64_t cab_grsd_cloud_labels(cab_ctx* ctx, float leaf, float r_normals, double rsd_radius_min, int32_t rsd_flags,
                              const float vp[3], int32_t* labels_plus1, int64_t cap);
int cab_grsd_cloud_set_labels(cab_ctx* ctx, const int32_t* labels_plus1, int64_t count);
int cab_grsd_cloud(cab_ctx* ctx, float leaf, float r_normals, double rsd_radius_min, int32_t rsd_flags, const float vp[3],
                   int32_t* hist21);

/* The colour half of VOSCH (extractVOSCH, grsd_colorCHLAC_tools.hpp:832-843: the 20 GRSD-21 bins followed by 117 colour
 * bins): rotation-invariant C3-HLAC (c3 = 1, pcl::C3HLAC_RI_Estimation, extractC3HLACSignature117 :787-812) or Color-CHLAC
 * (c3 = 0, pcl::ColorCHLAC_RI_Estimation, the variant the reference's shipped *_GRSD_CCHLAC.pcd vectors were made with),
 * color_chlac/include/color_chlac/color_chlac.hpp:1471-1528,1565-1782, from the voxels and leaf layouts the last
 * cab_grsd_batch left on the device.  rgb: one packed 0x00RRGGBB per point of that batch (the bits of PCL's float rgb
 * field); a voxel's colour is the truncated mean of its points' channels (pcl::VoxelGrid); thR/thG/thB binarise it
 * (value > threshold, the reference passes 127).  Subdivisions, hist_offsets, subdiv_b and the return value as in
 * cab_grsd_signatures.  hist: total x 117 floats, accumulated in the reference's order (the float sums of integer colour
 * products pass 2^24, so the order is part of the result) and normalised (:1764-1782). */
int64_t cab_color_chlac(cab_ctx* ctx, const uint32_t* rgb, int32_t c3, int32_t thR, int32_t thG, int32_t thB,
                        int32_t subdivision_size, int32_t off_x, int32_t off_y, int32_t off_z, int64_t* hist_offsets,
                        int32_t* subdiv_b, float* hist, int64_t cap);

/* ---- SVM classification of the signatures ---------------------------------------------
 * The consumer of GRSD in the reference pipeline (table_memory_grsd.cpp:1000-1020): replaces, per
 * feature vector, scaleFeature + svm_predict of cloud_algos::SVMClassification::process
 * (svm_classification.cpp:134-155, svm_classification.h:68-86) for a libsvm C-SVC / RBF model
 * (cloud_algos/svm/grsd_ijrr.model and the like).  The caller parses the model file
 * (svm_load_model) and the svm-scale range file (parseScaleParameterFile, svm_classification.h:129-185)
 * and hands over plain arrays:
 *   labels[nr_class], nr_sv[nr_class], rho[nr_class*(nr_class-1)/2] in libsvm's pair order,
 *   sv_coef[(nr_class-1) x total_sv], sv[total_sv x dim] dense (omitted sparse entries = 0).
 * fmin == fmax == NULL switches scaling off (scale_self_ = scale_file_ = false). */
int cab_svm_set_model(cab_ctx* ctx, int32_t dim, int32_t nr_class, int32_t total_sv, double gamma,
                      const int32_t* labels, const int32_t* nr_sv, const double* rho,
                      const double* sv_coef, const double* sv);
int cab_svm_set_scaling(cab_ctx* ctx, int32_t dim, double lower, double upper, const double* fmin,
                        const double* fmax);
/* features: n x dim floats (the f1..f<dim> channels, point-major).  point_class: n predicted labels
 * (the point_class channel, svm_classification.cpp:128,150).  dec_values (may be NULL):
 * n x nr_class*(nr_class-1)/2 decision values. */
int cab_svm_predict(cab_ctx* ctx, const float* features, int64_t n, int32_t dim, float* point_class,
                    double* dec_values);
/* Classifies the GRSD-21 histograms the last cab_grsd_batch left on the device (one per cluster;
 * f_i = (float)count_i as the GlobalRSD plugin emits them): cluster -> class without a host round
 * trip of the features.  point_class: nclusters floats. */
int cab_svm_predict_grsd(cab_ctx* ctx, float* point_class);

/* ---- statistical outlier removal (next row: cloud_algos/StatisticalNoiseRemoval) ----------
 * Replaces the kd-tree k-NN loop and the statistics of StatisticalNoiseRemoval::process
 * (cloud_algos/src/noise_removal.cpp:84-136) on the uploaded cloud:
 *   avg[cp] = mean distance from cp to its k - 1 nearest neighbours (k = neighborhood_size_ counts cp
 *             itself, which is skipped, :102-111; fp32 sqrt of the fp32 d2, summed in fp64), k-NN under
 *             the documented d2 rule, ties by index;
 *   mean / stddev of avg over the cloud, fp64, summed in input order (:112-121);
 *   keep[cp] = |avg[cp] - mean| < alpha * stddev (:131).
 * Non-finite points get avg = NaN, stay out of the statistics and are never kept.  cell_hint > 0 sets
 * the edge of the first search grid (a guess of the k-th neighbour distance), <= 0 derives it from the
 * cloud; the edge doubles until every query has k points within one edge.  avg / keep / mean / stddev
 * may be NULL.  cab_statistical_outliers returns the number of kept points or < 0; errors mirror the
 * plugin's checks (k < 2, alpha < 0, fewer than k points: noise_removal.cpp:51-62). */
int cab_knn_mean_distance(cab_ctx* ctx, int32_t k, float cell_hint, double* avg);
int64_t cab_statistical_outliers(cab_ctx* ctx, int32_t k, double alpha, float cell_hint, uint8_t* keep,
                                 double* avg, double* mean_out, double* stddev_out);

/* k-nearest-neighbour normals: the other normal estimation of the reference, nearestKSearch (i, k_) +
 * cloud_geometry::nearest::computePointNormal + flipNormalTowardsViewpoint [point_cloud_mapping, external] as in
 * TableObjectDetector::estimatePointNormals (cloud_tools/src/table_object_detector_passive.cpp:668-714, k_ = 10 at :169)
 * and CylinderEstimation::estimatePointNormals (cloud_algos/src/cylinder_fit_algo.cpp:138-203).  Same PCA, curvature and
 * flip as cab_normals over the k nearest points (the query included, ties by index) instead of a radius.  cell_hint as in
 * cab_knn_mean_distance.  nxyz_curv: n x 4 floats in input order (NaN for non-finite points), may be NULL.  Errors:
 * k < 3, fewer than k finite points.  The result is not kept on the device: cab_set_normals feeds it to cab_rsd. */
int cab_normals_knn(cab_ctx* ctx, int32_t k, const float vp[3], float cell_hint, float* nxyz_curv);

/* ---- Euclidean clustering (next row: the segmentation step that produces GRSD's object clusters) ----
 * Replaces cloud_geometry::nearest::extractEuclideanClusters(points, indices, tolerance, clusters, -1, -1, -1, -1,
 * min_pts) [point_cloud_mapping, external] as called at cloud_tools/src/table_object_detector_passive.cpp:293,567
 * and cloud_tools/src/table_object_detector_sr.cpp:370 (object_cluster_dist_tolerance 0.05, object_cluster_min_pts 30,
 * :181-182), on the uploaded cloud (the caller uploads the `indices` subset): connected components of the graph
 * "d2 <= tolerance^2" (the documented d2 rule), numbered in the order the reference's seed loop finds them (by
 * smallest index); components with fewer than min_pts (or, if max_pts > 0, more than max_pts) points are dropped.
 * The variant with a normal-angle test (nx_idx >= 0) is not implemented.  Builds its own grid (cell = tolerance),
 * clusters the whole cloud on this GPU whatever cab_set_shard says.
 * labels[n] (input order, may be NULL): cluster id, -1 for dropped components and non-finite points.
 * Returns the number of clusters or < 0. */
int64_t cab_euclidean_clusters(cab_ctx* ctx, double tolerance, int32_t min_pts, int32_t max_pts, int32_t* labels);
/* Host helper: labels -> the reference's vector<vector<int>> as CSR.  offsets[n_clusters + 1]; indices (may be NULL)
 * holds every cluster's point indices in ascending order, cluster after cluster -- gather the points with it and hand
 * them, with offsets, to cab_upload_clusters / cab_grsd_batch.  Returns the number of labelled points or < 0. */
int64_t cab_cluster_csr(const int32_t* labels, int64_t n, int32_t n_clusters, int32_t* offsets, int32_t* indices);

/* The table-plane step in front of the object clustering: fitSACPlane (cloud_tools/src/table_object_detector_passive.cpp:
 * 621-659, table_object_detector_sr.cpp): sample_consensus::MSAC over SACModelPlane (setMaxIterations(500),
 * setProbability(0.99)) [point_cloud_mapping, external], computeCoefficients, refineCoefficients (least squares over the
 * inliers), selectWithinDistance(coeff, threshold) and projectPointsInPlace.
 * xyz: n host points (stride floats each); indices: optional list of n_idx point indices the model is fitted to (the
 * detector passes one cluster of Z-parallel points), NULL = all points.  triples: n_triples x 3 POSITIONS into that list
 * (or into the cloud), the sample sequence -- the caller's, because the reference's rand() stream is not reproducible;
 * a hypothesis is the plane through the next triple; a degenerate triple (repeated or collinear points) is skipped, as
 * upstream getSamples draws again.  The loop runs while iterations < k (k = log(1 - probability) / log(1 - w^3) after
 * every improvement, w = inlier ratio of the best model), for at most max_iterations + 1 hypotheses and n_triples
 * samples, exactly as the sequential loop would -- the device scores 32 hypotheses per launch.
 * coeff: a, b, c, d of the refined plane (unit normal on the sampled model's side).  inliers: the indices within
 * `threshold` of the refined plane, in list order; projected_xyz: their projections (3 floats each); both optional, room
 * for `cap` entries.  iterations_run / best_iteration (optional): hypotheses scored, position of the winning triple in
 * the sample sequence (-1: none valid).  Returns the number of inliers (0 with coeff = 0 if no model was found) or < 0. */
int64_t cab_fit_plane_msac(cab_ctx* ctx, const float* xyz, int64_t n, int32_t stride, const int32_t* indices, int64_t n_idx,
                           double threshold, int32_t max_iterations, double probability, const int32_t* triples,
                           int64_t n_triples, double coeff[4], int32_t* inliers, float* projected_xyz, int64_t cap,
                           int32_t* iterations_run, int32_t* best_iteration);

/* ---- point feature histograms (next row: cloud_algos/PointFeatureHistogram) ----------------
 * Replaces the hot loops of PointFeatureHistogram::process (cloud_algos/src/pfh.cpp:181-350; pair features
 * cloud_algos/include/cloud_algos/pfh.h:102-238) on the uploaded cloud with its normals (cab_set_normals or
 * cab_normals; the grid must cover `radius`): for every point the "star" pair features alpha, beta, gamma
 * [, delta] towards its neighbours within `radius` (<= max_nn nearest, max_nn <= 0: all), one histogram of
 * `quantum` bins per feature with increments of 100 / k, then the options of the plugin:
 *   CAB_PFH_AVERAGE       1/d2-weighted average of the neighbours' histograms = FPFH (average_, :303-333)
 *   CAB_PFH_DIFFERENTIAL  bin-to-bin differences (differential_, :337-350)
 *   CAB_PFH_CHECK_FLIP / CAB_PFH_ABS_ANGLES / CAB_PFH_USE_DIST  check_flip_, abs_angles_, use_dist_
 *   CAB_PFH_COMBINE       combine_: ONE histogram of quantum ^ (3 or 4) bins per point, the feature indices being the
 *                         digits of the bin number in the reference's order (:113-121, 239-258); an invalid pair adds
 *                         100 / k / nbins to every bin (:279-281); no differences (:345).  At most 4096 bins.
 * The plugin's defaults (pfh.h:83-93) are radius 0.03, max_nn 100, quantum 9, CHECK_FLIP | AVERAGE.  hist: n x nbins
 * floats, nbins = quantum * (3 or 4), or quantum ^ (3 or 4) with COMBINE, point-major, input order (channel f<b+1> of
 * point i is hist[i * nbins + b]). */
#define CAB_PFH_USE_DIST 1
#define CAB_PFH_DIFFERENTIAL 2
#define CAB_PFH_CHECK_FLIP 4
#define CAB_PFH_ABS_ANGLES 8
#define CAB_PFH_AVERAGE 16
#define CAB_PFH_COMBINE 32 /* combine_: one n-D histogram of quantum ^ (3 or 4) bins per point (pfh.cpp:47-57, 239-258) */
int cab_pfh(cab_ctx* ctx, double radius, int32_t max_nn, int32_t quantum, int32_t flags, float* hist);

/* ---- device plumbing (bench / multi-GPU) ---------------------------------------------- */
#define CAB_BUF_POS_SORTED 0  /* float4[n]  x,y,z,0 in sorted order */
#define CAB_BUF_NRM_SORTED 1  /* float4[n]  nx,ny,nz,curvature in sorted order */
#define CAB_BUF_RSD_SORTED 2  /* float2[n]  r_min,r_max in sorted order */
#define CAB_BUF_PERM 3        /* int32[n]   sorted position -> input index */
void* cab_device_ptr(cab_ctx* ctx, int32_t which);
void* cab_stream(cab_ctx* ctx); /* cudaStream_t all work of this context is enqueued on */
/* Copy device results (sorted order) to host arrays in input order. Any pointer may be NULL. */
int cab_download(cab_ctx* ctx, float* nxyz_curv, float* r_min, float* r_max);

/* r_dif of the last cab_rsd in input order: (float)(max_radius - min_radius) with the subtraction in double, before the
 * radii are rounded to float -- the value LocalRadiusEstimation stores in its r_dif channel (radius_estimation.cpp:206);
 * r_max - r_min of the rounded outputs can differ from it in the last bit. */
int cab_download_rdif(cab_ctx* ctx, float* r_dif);

/* Copy the sorted-order slice [begin, end) of the results to host: normals (4 floats per point),
 * radii (2 floats: r_min, r_max) and the sorted-position -> input-index map.  Any pointer may be
 * NULL.  A rank of a sharded run downloads its own cab_shard_range this way. */
int cab_download_sorted(cab_ctx* ctx, int64_t begin, int64_t end, float* nxyz_curv, float* rmin_rmax,
                        int32_t* input_index);

/* ---- one step, one call ------------------------------------------------------------------
 * cab_build_grid(cell) + cab_normals((float)r) + cab_rsd(r) on the cloud last given to cab_upload_cloud /
 * cab_set_cloud_device / cab_comm_upload_cloud, enqueued back to back with one host synchronisation at the end instead
 * of one per stage (the NormalEstimation -> LocalRadiusEstimation chain of cloud_algos/sample_pipeline.yaml as a frame
 * loop would run it).  Results stay on the device (cab_download*, CAB_BUF_*).  When the context belongs to a group
 * (cab_comm_init*), the step also concatenates the ranks' results: the RSD kernel stores every packet's normals and
 * radii into peer memory over NVLink -- into every rank's copy of the concatenated arrays (replicated layout: when the
 * call returns EVERY rank holds the results of ALL queries) or into the arrays of the rank that owns the point's input
 * index (input-range layout, cab_comm_set_layout); see cab_comm_device_ptr, cab_comm_download_range. */
int cab_step_normals_rsd(cab_ctx* ctx, float cell, double r, int32_t max_nn_normals, const float vp[3], int32_t max_nn_rsd,
                         int32_t ndiv, double plane_radius, int32_t flags);

/* ---- multi-GPU groups ---------------------------------------------------------------------
 * One cab_ctx per GPU; the contexts of a group run the same calls in the same order (SPMD), each from its own host
 * thread or process.  Joining a group also sets the shard (cab_set_shard(rank, world)).
 * One GPU per rank is the intended layout.  Ranks that share a GPU (tests) need CUDA_MODULE_LOADING=EAGER and
 * CUDA_DEVICE_MAX_CONNECTIONS >= 2 x (ranks on the GPU) in the environment before CUDA is initialised: a step's kernels
 * wait for the peers' flags, and a lazily loaded kernel's first launch, or a kernel queued behind a waiting kernel in a
 * shared hardware queue, would stall the step until the timeout; connecting such a group without them fails with
 * CAB_ERR_STATE.
 *   cab_comm_init_local   contexts of ONE process (e.g. a plugin host with one worker thread per GPU): peer access
 *                         between the devices, nothing else needed.
 *   cab_comm_init         one process per GPU (the torchrun / MPI layout): rank 0 obtains an id with cab_comm_get_id
 *                         (ncclGetUniqueId; libnccl.so.2 is loaded on first use), the application hands it to every
 *                         rank, every rank calls cab_comm_init.  NCCL carries the bootstrap (CUDA IPC handles of the
 *                         exchange buffers) and cab_comm_allreduce_i32; the data path uses peer memory directly.
 *   cab_comm_reserve / cab_comm_connect   the same without NCCL: every rank reserves buffers for clouds of up to
 *                         max_points points and gets a blob of CAB_COMM_BLOB_BYTES; the application gathers the blobs
 *                         of all ranks (rank order) by whatever means it has and gives them to cab_comm_connect. */
#define CAB_COMM_ID_BYTES 128
#define CAB_COMM_BLOB_BYTES 512
int cab_comm_get_id(char id[CAB_COMM_ID_BYTES]);
int cab_comm_init(cab_ctx* ctx, const char id[CAB_COMM_ID_BYTES], int32_t rank, int32_t world);
int cab_comm_init_local(cab_ctx** ctxs, int32_t world);
int cab_comm_reserve(cab_ctx* ctx, int32_t rank, int32_t world, int64_t max_points, void* blob);
int cab_comm_connect(cab_ctx* ctx, const void* blobs);
int cab_comm_free(cab_ctx* ctx);
/* Where the step's concatenated results live (the same choice on every rank; default REPLICATED):
 *   CAB_COMM_LAYOUT_REPLICATED    every rank ends up holding the results of ALL queries, slab after slab in sorted order
 *                                 plus the input index of every entry (cab_comm_device_ptr CAB_BUF_*_SORTED / _PERM);
 *                                 cab_comm_download_range serves any input range on any rank.  world x the traffic.
 *   CAB_COMM_LAYOUT_INPUT_RANGES  rank g ends up holding the results of input indices [n g / world, n (g + 1) / world) in
 *                                 INPUT order (the ranks' arrays, rank after rank, are the channels radius_estimation.cpp:
 *                                 204-214 appends): every result crosses NVLink once, to its owner;
 *                                 cab_comm_download_range serves ranges inside the rank's own range
 *                                 (cab_comm_device_ptr CAB_BUF_NRM_INPUT_RANGE / CAB_BUF_RSD_INPUT_RANGE). */
#define CAB_COMM_LAYOUT_REPLICATED 0
#define CAB_COMM_LAYOUT_INPUT_RANGES 1
#define CAB_BUF_NRM_INPUT_RANGE 4 /* float4[hi-lo] nx,ny,nz,curvature of this rank's input range, input order */
#define CAB_BUF_RSD_INPUT_RANGE 5 /* float2[hi-lo] r_min,r_max of this rank's input range, input order */
int cab_comm_set_layout(cab_ctx* ctx, int32_t layout);
/* Shard balance of a group.  The rows are cut by a cost model (sampled cell histogram -> candidates per row); with the
 * feedback on (default) every step also leaves each rank's measured normals + RSD time with every rank, and the next
 * step's cuts give a rank that took longer than the mean a smaller share (steps whose passes take less than half a
 * millisecond leave the shares alone: that is latency, not work).  Results never depend on the cuts; all ranks
 * must use the same setting (a step fails with CAB_ERR_STATE if the ranks' cuts disagree).  Setting it resets the
 * shares to equal. */
int cab_comm_set_feedback(cab_ctx* ctx, int32_t on);
/* The ranks' shares of the modelled cost, set by hand (unequal GPUs, or a test that wants the cuts elsewhere): `count` =
 * world positive numbers, normalised to sum 1; the same numbers on every rank.  The feedback, if on, continues from them. */
int cab_comm_set_shares(cab_ctx* ctx, const double* shares, int32_t count);
/* Replicates a host cloud on every rank of the group: xyz points at the WHOLE cloud (n points, packed xyz) on every rank;
 * a rank uploads rows [n*rank/world, n*(rank+1)/world) over its own PCIe link and copies them into the peers' buffers
 * over NVLink (copy engines), then waits for the other slices.  Without a group: cab_upload_cloud. */
int cab_comm_upload_cloud(cab_ctx* ctx, const float* xyz, int64_t n, int32_t stride);
/* Results of the last cab_step_normals_rsd of the group for the input-order range [j0, j1) -- any range, on any rank
 * (replicated layout), or any range inside [n rank / world, n (rank + 1) / world) (input-range layout):
 * nxyz_curv (j1-j0) x 4, r_min / r_max (j1-j0) floats; any pointer may be NULL.  With j0 = n*rank/world, j1 =
 * n*(rank+1)/world and pointers into ONE host array shared by the ranks, the ranks together leave the channels
 * LocalRadiusEstimation appends (radius_estimation.cpp:204-214) in input order, each paying 1/world of the copy. */
int cab_comm_download_range(cab_ctx* ctx, int64_t j0, int64_t j1, float* nxyz_curv, float* r_min, float* r_max);
/* The concatenated arrays of the last exchanged step on this rank: CAB_BUF_NRM_SORTED / _RSD_SORTED / _PERM, *count
 * entries (all finite points of the cloud, slab after slab). */
void* cab_comm_device_ptr(cab_ctx* ctx, int32_t which, int64_t* count);
/* In-place sum over the group of `count` int32 values in host memory: the integer histograms of GRSD batches sharded by
 * cluster, or of one large cloud sharded by voxels (one small ncclAllReduce; bit-exact in any order). */
int cab_comm_allreduce_i32(cab_ctx* ctx, int32_t* values, int64_t count);

int cab_profile(const cab_ctx* ctx, cab_timings* out);
int cab_version(void);

#ifdef __cplusplus
}
#endif
#endif
