// plugin_capi.cpp -- a flat C wrapper around the C++ plugins so that the Python tests (ctypes)
// can drive them exactly like the reference's callers do: look the class up by its pluginlib name,
// init(nh), then per message pre() -> (set public fields) -> process() -> output() -> post()
// (cloud_algos.h:79-97 of the reference; dyn_obj_store/src/table_memory_grsd.cpp:683-705,974-994).
#include <cstring>
#include <string>

#include <pluginlib/class_loader.h>
#include <cloud_algos/cloud_algos.h>
#include <cloud_algos/normal_estimation.h>
#include <cloud_algos/radius_estimation.h>
#include <cloud_algos/global_rsd.h>
#include <cloud_algos/svm_classification.h>
#include <cloud_algos/noise_removal.h>
#include <cloud_algos/pfh.h>
#include <cloud_algos/pcd_io.h>
#include <point_cloud_mapping/geometry/nearest.h>
#include <cloud_tools/fit_sac_plane.h>

using namespace cloud_algos;

namespace {
struct Handle {
  std::string name;
  CloudAlgo* algo = nullptr;
  ros::NodeHandle nh{"~"};
  ros::Publisher pub;
  boost::shared_ptr<const sensor_msgs::PointCloud> out, aux;
  std::string result;
};
}  // namespace

extern "C" {

void* capi_create(const char* lookup_name) {
  static pluginlib::ClassLoader<CloudAlgo> loader("cloud_algos", "cloud_algos::CloudAlgo");
  try {
    loader.loadLibraryForClass(lookup_name);
    Handle* h = new Handle();
    h->name = lookup_name;
    h->algo = loader.createClassInstance(lookup_name);
    h->algo->init(h->nh);
    h->pub = h->algo->createPublisher(h->nh);
    return h;
  } catch (pluginlib::PluginlibException&) {
    return nullptr;
  }
}

void capi_destroy(void* hv) {
  Handle* h = (Handle*)hv;
  if (!h) return;
  delete h->algo;
  delete h;
}

void capi_set_param(void* hv, const char* key, double value) { ((Handle*)hv)->nh.setParam(key, value); }
void capi_set_param_str(void* hv, const char* key, const char* value) { ((Handle*)hv)->nh.setParam(key, std::string(value)); }
int capi_set_field_str(void* hv, const char* field, const char* value) {
  const std::string f(field);
  if (SVMClassification* a = dynamic_cast<SVMClassification*>(((Handle*)hv)->algo)) {
    if (f == "model_file_name_") a->model_file_name_ = value; else if (f == "scale_file_name_") a->scale_file_name_ = value;
    else return -1;
    return 0;
  }
  return -1;
}

// Public-field assignment after pre(), the way table_memory_grsd.cpp:975-981 pokes GlobalRSD.
int capi_set_field(void* hv, const char* field, double value) {
  Handle* h = (Handle*)hv;
  const std::string f(field);
  if (f == "verbosity_level_") { h->algo->verbosity_level_ = (int)value; return 0; }
  if (LocalRadiusEstimation* a = dynamic_cast<LocalRadiusEstimation*>(h->algo)) {
    if (f == "radius_") a->radius_ = value; else if (f == "max_nn_") a->max_nn_ = (int)value;
    else if (f == "plane_radius_") a->plane_radius_ = value; else if (f == "distance_div_") a->distance_div_ = (int)value;
    else if (f == "point_label_") a->point_label_ = (int)value; else if (f == "rmin2curvature_") a->rmin2curvature_ = value != 0;
    else return -1;
    return 0;
  }
  if (NormalEstimation* a = dynamic_cast<NormalEstimation*>(h->algo)) {
    if (f == "radius_") a->radius_ = value; else if (f == "max_nn_") a->max_nn_ = (int)value;
    else if (f == "vp_x_") a->vp_x_ = value; else if (f == "vp_y_") a->vp_y_ = value; else if (f == "vp_z_") a->vp_z_ = value;
    else return -1;
    return 0;
  }
  if (SVMClassification* a = dynamic_cast<SVMClassification*>(h->algo)) {
    if (f == "scale_self_") a->scale_self_ = value != 0; else if (f == "scale_file_") a->scale_file_ = value != 0;
    else return -1;
    return 0;
  }
  if (PointFeatureHistogram* a = dynamic_cast<PointFeatureHistogram*>(h->algo)) {
    if (f == "radius_") a->radius_ = value; else if (f == "max_nn_") a->max_nn_ = (int)value;
    else if (f == "quantum_") a->quantum_ = (int)value; else if (f == "use_dist_") a->use_dist_ = value != 0;
    else if (f == "combine_") a->combine_ = value != 0; else if (f == "differential_") a->differential_ = value != 0;
    else if (f == "check_flip_") a->check_flip_ = value != 0; else if (f == "abs_angles_") a->abs_angles_ = value != 0;
    else if (f == "average_") a->average_ = value != 0; else if (f == "point_label_") a->point_label_ = (int)value;
    else return -1;
    return 0;
  }
  if (StatisticalNoiseRemoval* a = dynamic_cast<StatisticalNoiseRemoval*>(h->algo)) {
    if (f == "alpha_") a->alpha_ = value; else if (f == "neighborhood_size_") a->neighborhood_size_ = (int)value;
    else if (f == "min_nr_pts_") a->min_nr_pts_ = (int)value;
    else return -1;
    return 0;
  }
  if (GlobalRSD* a = dynamic_cast<GlobalRSD*>(h->algo)) {
    if (f == "width_") a->width_ = value; else if (f == "step_") a->step_ = (int)value;
    else if (f == "min_voxel_pts_") a->min_voxel_pts_ = (int)value; else if (f == "label_") a->label_ = (int)value;
    else if (f == "publish_cloud_centroids_") a->publish_cloud_centroids_ = value != 0;
    else if (f == "publish_cloud_vrsd_") a->publish_cloud_vrsd_ = value != 0;
    else return -1;
    return 0;
  }
  return -1;
}

void capi_pre(void* hv) { ((Handle*)hv)->algo->pre(); }
void capi_post(void* hv) { ((Handle*)hv)->algo->post(); }

// process(): xyz is n x 3, channel c has `names[c]` and n values at values[c].
const char* capi_process(void* hv, const float* xyz, int n, int nchan, const char* const* names, const float* const* values) {
  Handle* h = (Handle*)hv;
  boost::shared_ptr<sensor_msgs::PointCloud> in(new sensor_msgs::PointCloud());
  in->header.frame_id = "base_link";
  in->points.resize(n);
  for (int i = 0; i < n; ++i) {
    in->points[i].x = xyz[3 * i];
    in->points[i].y = xyz[3 * i + 1];
    in->points[i].z = xyz[3 * i + 2];
  }
  in->channels.resize(nchan);
  for (int c = 0; c < nchan; ++c) {
    in->channels[c].name = names[c];
    in->channels[c].values.assign(values[c], values[c] + n);
  }
  boost::shared_ptr<const sensor_msgs::PointCloud> cin = in;
  h->out.reset();
  if (LocalRadiusEstimation* a = dynamic_cast<LocalRadiusEstimation*>(h->algo)) {
    h->result = a->process(cin);
    if (a->output_valid_) h->out = a->output();
  } else if (NormalEstimation* a = dynamic_cast<NormalEstimation*>(h->algo)) {
    h->result = a->process(cin);
    if (a->output_valid_) h->out = a->output();
  } else if (GlobalRSD* a = dynamic_cast<GlobalRSD*>(h->algo)) {
    h->result = a->process(cin);
    if (a->output_valid_) h->out = a->output();
    h->aux = a->getVRSD();
  } else if (SVMClassification* a = dynamic_cast<SVMClassification*>(h->algo)) {
    h->result = a->process(cin);
    if (a->output_valid_) h->out = a->output();
  } else if (StatisticalNoiseRemoval* a = dynamic_cast<StatisticalNoiseRemoval*>(h->algo)) {
    h->result = a->process(cin);
    if (a->output_valid_) h->out = a->output();
  } else if (PointFeatureHistogram* a = dynamic_cast<PointFeatureHistogram*>(h->algo)) {
    h->result = a->process(cin);
    if (a->output_valid_)  // output() returns the cloud by value (pfh.h:62)
      h->out = boost::shared_ptr<const sensor_msgs::PointCloud>(new sensor_msgs::PointCloud(a->output()));
  } else {
    h->result = "unknown plugin type";
  }
  if (h->algo->output_valid_ && h->out) h->pub.publish(h->out);
  return h->result.c_str();
}

// The reference's stand-alone node path (cloud_algos.h:46-117): a CloudAlgoNode subscribes the algorithm to its default
// input topic, a message on that topic runs pre / process / publish(output) / post.  Returns the messages published.
int capi_radius_node_roundtrip(const float* xyz, const float* nx, const float* ny, const float* nz, int n, double radius) {
  boost::shared_ptr<sensor_msgs::PointCloud> cloud(new sensor_msgs::PointCloud());
  cloud->points.resize((size_t)n);
  for (int i = 0; i < n; ++i) {
    cloud->points[i].x = xyz[3 * i];
    cloud->points[i].y = xyz[3 * i + 1];
    cloud->points[i].z = xyz[3 * i + 2];
  }
  if (nx) {
    cloud->channels.resize(3);
    const char* names[3] = {"nx", "ny", "nz"};
    const float* src[3] = {nx, ny, nz};
    for (int c = 0; c < 3; ++c) {
      cloud->channels[c].name = names[c];
      cloud->channels[c].values.assign(src[c], src[c] + n);
    }
  }
  ros::inject<sensor_msgs::PointCloud>(LocalRadiusEstimation::default_input_topic(), cloud);
  LocalRadiusEstimation a;
  ros::NodeHandle nh("~");
  nh.setParam("radius", radius);
  CloudAlgoNode<LocalRadiusEstimation> node(nh, a);
  ros::spin();
  return node.pub_.getNumPublished();
}

int capi_output_valid(void* hv) { return ((Handle*)hv)->algo->output_valid_ ? 1 : 0; }
int capi_num_published(void* hv) { return ((Handle*)hv)->pub.getNumPublished(); }
const char* capi_topic(void* hv) { return ((Handle*)hv)->pub.getTopic().c_str(); }

static const sensor_msgs::PointCloud* pick(Handle* h, int which) { return which == 0 ? h->out.get() : h->aux.get(); }
int capi_out_size(void* hv, int which) { const sensor_msgs::PointCloud* c = pick((Handle*)hv, which); return c ? (int)c->points.size() : -1; }
int capi_out_num_channels(void* hv, int which) { const sensor_msgs::PointCloud* c = pick((Handle*)hv, which); return c ? (int)c->channels.size() : -1; }
const char* capi_out_channel_name(void* hv, int which, int c) { return pick((Handle*)hv, which)->channels[c].name.c_str(); }
const float* capi_out_channel(void* hv, int which, int c) { return pick((Handle*)hv, which)->channels[c].values.data(); }
const float* capi_out_points(void* hv, int which) { const sensor_msgs::PointCloud* c = pick((Handle*)hv, which); return c->points.empty() ? nullptr : &c->points[0].x; }

// PCD I/O helpers (host/include/cloud_algos/pcd_io.h): returns the number of points, -1 on failure; xyz / nrm hold
// up to cap points (nrm may be NULL); *has_normals tells whether the file carried normal_x/y/z.
int capi_pcd_read(const char* name, float* xyz, float* nrm, int cap, int* has_normals) {
  std::vector<float> p, nn;
  if (!readPCDXYZ(name, p, &nn)) return -1;
  const int n = (int)(p.size() / 3);
  if (has_normals) *has_normals = nn.empty() ? 0 : 1;
  for (int i = 0; i < n && i < cap; ++i)
    for (int a = 0; a < 3; ++a) {
      if (xyz) xyz[3 * i + a] = p[3 * i + a];
      if (nrm && !nn.empty()) nrm[3 * i + a] = nn[3 * i + a];
    }
  return n;
}
// packed colours of a PCD file (0x00RRGGBB per point): returns the number of points, 0 if the file has no rgb field, -1 on failure
int capi_pcd_read_rgb(const char* name, unsigned* rgb, int cap) {
  std::vector<float> p;
  std::vector<uint32_t> c;
  if (!readPCDXYZ(name, p, 0, 0, &c)) return -1;
  for (size_t i = 0; i < c.size() && (int)i < cap; ++i) rgb[i] = c[i];
  return (int)c.size();
}
int capi_write_feature(const char* name, const float* data, int hist_num, int dim, int remove_0) {
  std::vector<std::vector<float> > f(hist_num, std::vector<float>(dim));
  for (int h = 0; h < hist_num; ++h) f[h].assign(data + (size_t)h * dim, data + (size_t)(h + 1) * dim);
  return writeFeature(name, f, remove_0 != 0) ? 0 : -1;
}

// cloud_geometry::nearest::extractEuclideanClusters through its reference signature; the clusters come back as
// cluster_of[k] for the k-th entry of indices (-1: in no cluster) plus the flattened members in cluster order.
// Returns the number of clusters, -1 if the call reported an error.
int capi_extract_euclidean_clusters(const float* xyz, int n, const int* indices, int n_idx, double tolerance, int nx_idx,
                                    unsigned min_pts, int* cluster_of, int* flat_members) {
  sensor_msgs::PointCloud cloud;
  cloud.points.resize(n);
  for (int i = 0; i < n; ++i) {
    cloud.points[i].x = xyz[3 * i];
    cloud.points[i].y = xyz[3 * i + 1];
    cloud.points[i].z = xyz[3 * i + 2];
  }
  std::vector<int> idx(indices, indices + n_idx);
  std::vector<std::vector<int> > clusters;
  cloud_geometry::nearest::extractEuclideanClusters(cloud, idx, tolerance, clusters, nx_idx, nx_idx < 0 ? -1 : nx_idx + 1,
                                                    nx_idx < 0 ? -1 : nx_idx + 2, 0.0, min_pts);
  if (!cloud_geometry::nearest::lastEuclideanClusterError().empty()) return -1;
  std::vector<int> where(n, -1);
  for (int k = 0; k < n_idx; ++k) where[idx[k]] = k;
  for (int k = 0; k < n_idx; ++k) cluster_of[k] = -1;
  int w = 0;
  for (size_t c = 0; c < clusters.size(); ++c)
    for (int i : clusters[c]) {
      cluster_of[where[i]] = (int)c;
      flat_members[w++] = i;
    }
  return (int)clusters.size();
}

// GlobalRSD::process_batch: clusters given as one concatenated cloud + offsets, nx/ny/nz per point; out: nc x 21 floats.
// Returns 0, or -1 with the error string in the handle's result.
int capi_grsd_process_batch(void* hv, const float* xyz, const float* nxyz, const int* offsets, int nc, float* out) {
  Handle* h = (Handle*)hv;
  GlobalRSD* a = dynamic_cast<GlobalRSD*>(h->algo);
  if (!a) { h->result = "not a GlobalRSD"; return -1; }
  std::vector<boost::shared_ptr<const sensor_msgs::PointCloud> > clusters(nc), outs;
  for (int c = 0; c < nc; ++c) {
    boost::shared_ptr<sensor_msgs::PointCloud> in(new sensor_msgs::PointCloud());
    const int b = offsets[c], m = offsets[c + 1] - offsets[c];
    in->points.resize(m);
    in->channels.resize(3);
    in->channels[0].name = "nx"; in->channels[1].name = "ny"; in->channels[2].name = "nz";
    for (int k = 0; k < 3; ++k) in->channels[k].values.resize(m);
    for (int i = 0; i < m; ++i) {
      in->points[i].x = xyz[3 * (size_t)(b + i)];
      in->points[i].y = xyz[3 * (size_t)(b + i) + 1];
      in->points[i].z = xyz[3 * (size_t)(b + i) + 2];
      for (int k = 0; k < 3; ++k) in->channels[k].values[i] = nxyz[3 * (size_t)(b + i) + k];
    }
    clusters[c] = in;
  }
  h->result = a->process_batch(clusters, outs);
  if (h->result != "ok") return -1;
  for (int c = 0; c < nc; ++c)
    for (int i = 0; i < 21; ++i) out[21 * (size_t)c + i] = outs[c]->channels[i].values.at(0);
  return 0;
}

// cloud_tools::fitSACPlane through the member's signature: xyz is updated in place with the projected inliers.
// Returns the number of inliers (coeff filled), 0 if no model, -1 if there were too few indices, -2 on a library error.
int capi_fit_sac_plane(float* xyz, int n, const int* indices, int n_idx, double threshold, int min_pts, unsigned seed,
                       int* inliers, double* coeff) {
  sensor_msgs::PointCloud cloud;
  cloud.points.resize(n);
  for (int i = 0; i < n; ++i) {
    cloud.points[i].x = xyz[3 * i];
    cloud.points[i].y = xyz[3 * i + 1];
    cloud.points[i].z = xyz[3 * i + 2];
  }
  std::vector<int> idx(indices, indices + n_idx), in;
  std::vector<double> c;
  const int rc = cloud_tools::fitSACPlane(&cloud, &idx, in, c, threshold, min_pts, seed);
  if (!cloud_tools::lastFitSACPlaneError().empty()) return -2;
  if (rc < 0) return -1;
  for (size_t i = 0; i < in.size(); ++i) inliers[i] = in[i];
  for (size_t k = 0; k < c.size(); ++k) coeff[k] = c[k];
  for (int i = 0; i < n; ++i) {
    xyz[3 * i] = cloud.points[i].x;
    xyz[3 * i + 1] = cloud.points[i].y;
    xyz[3 * i + 2] = cloud.points[i].z;
  }
  return (int)in.size();
}

int capi_list_requires(void* hv, char* buf, int cap) {
  std::string s;
  for (auto& r : ((Handle*)hv)->algo->requires()) s += r + ",";
  s += "|";
  for (auto& p : ((Handle*)hv)->algo->provides()) s += p + ",";
  std::strncpy(buf, s.c_str(), cap - 1);
  buf[cap - 1] = 0;
  return (int)s.size();
}

}  // extern "C"
