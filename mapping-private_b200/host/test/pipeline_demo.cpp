// pipeline_demo.cpp -- drives the three plugins the way the reference's nodes / callers do:
//   NormalEstimation -> LocalRadiusEstimation (CloudAlgoNode::input_cb order: pre, process,
//   publish(output) if output_valid_, post) -> GlobalRSD (table_memory_grsd.cpp:974-996).
// Input: a synthetic sphere patch generated in-process; prints a few sanity numbers.
// Needs a B200 (there is no CPU fallback); exits 2 with the plugin's message otherwise.
#include <cmath>
#include <cstdio>
#include <random>

#include <pluginlib/class_loader.h>
#include <cloud_algos/cloud_algos.h>
#include <cloud_algos/normal_estimation.h>
#include <cloud_algos/radius_estimation.h>
#include <cloud_algos/global_rsd.h>

using namespace cloud_algos;

int main() {
  boost::shared_ptr<sensor_msgs::PointCloud> cloud(new sensor_msgs::PointCloud());
  std::mt19937 rng(7);
  std::normal_distribution<float> g(0.f, 1.f);
  const int n = 20000;
  cloud->points.resize(n);
  for (int i = 0; i < n; ++i) {
    float x = g(rng), y = g(rng), z = g(rng), l = std::sqrt(x * x + y * y + z * z);
    cloud->points[i].x = 0.5f + 0.05f * x / l;
    cloud->points[i].y = 0.5f + 0.05f * y / l;
    cloud->points[i].z = 1.0f + 0.05f * z / l;
  }
  ros::NodeHandle nh("~");
  nh.setParam("radius", 0.02);
  nh.setParam("max_nn", 0);

  pluginlib::ClassLoader<CloudAlgo> loader("cloud_algos", "cloud_algos::CloudAlgo");
  NormalEstimation* ne = (NormalEstimation*)loader.createClassInstance("cloud_algos/NormalEstimation");
  LocalRadiusEstimation* rsd = (LocalRadiusEstimation*)loader.createClassInstance("cloud_algos/LocalRadiusEstimation");
  GlobalRSD* grsd = (GlobalRSD*)loader.createClassInstance("cloud_algos/GlobalRSD");
  ne->init(nh);
  rsd->init(nh);
  grsd->init(nh);

  ne->pre();
  std::string r1 = ne->process(cloud);
  if (!ne->output_valid_) { std::fprintf(stderr, "NormalEstimation: %s\n", r1.c_str()); return 2; }
  boost::shared_ptr<const sensor_msgs::PointCloud> with_normals = ne->output();
  ne->post();

  rsd->pre();
  std::string r2 = rsd->process(with_normals);
  if (!rsd->output_valid_) { std::fprintf(stderr, "LocalRadiusEstimation: %s\n", r2.c_str()); return 2; }
  boost::shared_ptr<const sensor_msgs::PointCloud> radii = rsd->output();
  rsd->post();

  grsd->pre();
  grsd->min_voxel_pts_ = 0;
  grsd->publish_cloud_vrsd_ = true;
  std::string r3 = grsd->process(with_normals);
  if (!grsd->output_valid_) { std::fprintf(stderr, "GlobalRSD: %s\n", r3.c_str()); return 2; }
  boost::shared_ptr<const sensor_msgs::PointCloud> hist = grsd->output();
  grsd->post();

  const int ri = getChannelIndex(radii, "r_min");
  double mean_rmin = 0;
  for (int i = 0; i < n; ++i) mean_rmin += radii->channels[ri].values[i];
  std::printf("%s %s %s | mean r_min %.4f (sphere R = 0.05) | GRSD:", r1.c_str(), r2.c_str(), r3.c_str(), mean_rmin / n);
  for (size_t c = 0; c < hist->channels.size(); ++c) std::printf(" %s=%g", hist->channels[c].name.c_str(), hist->channels[c].values.at(0));
  std::printf("\n");
  delete ne;
  delete rsd;
  delete grsd;
  return 0;
}
