// svm_classification.cpp -- cloud_algos::SVMClassification on the B200.
// Contract of cloud_algos/src/svm_classification.cpp of the reference: rosparams read in pre()
// (:12-20), requires f1 (:27-33), provides point_class (:35-40); process() finds the consecutive
// f1..fN channels (:44-66), loads the model (:80-90, "incorrect model file"), sets up scaling
// (:92-117, "incorrect scale parameter file"), copies the cloud and appends point_class (:123-133),
// classifies every point (:135-155) and reports the accuracy against point_label if present.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include <cloud_algos/cloud_algos.h>
#include <cloud_algos/svm_classification.h>

#include "cloud_algos_b200.h"

using namespace cloud_algos;

bool SvmModelData::load (const char* file_name, int min_dim)
{
  std::ifstream fs (file_name);
  if (!fs.is_open ()) return false;
  std::string line, svm_type, kernel_type;
  int nr_class = 0;
  bool in_sv = false;
  while (std::getline (fs, line))
  {
    std::istringstream is (line);
    std::string key;
    if (!(is >> key)) continue;
    if (key == "SV") { in_sv = true; break; }
    if (key == "svm_type") is >> svm_type;
    else if (key == "kernel_type") is >> kernel_type;
    else if (key == "gamma") is >> gamma;
    else if (key == "nr_class") is >> nr_class;
    else if (key == "total_sv") is >> total_sv;
    else if (key == "rho") { double v; while (is >> v) rho.push_back (v); }
    else if (key == "label") { int v; while (is >> v) labels.push_back (v); }
    else if (key == "nr_sv") { int v; while (is >> v) nr_sv.push_back (v); }
  }
  if (!in_sv || svm_type != "c_svc" || kernel_type != "rbf" || nr_class < 2 || total_sv < 1) return false;
  if ((int) labels.size () != nr_class || (int) nr_sv.size () != nr_class ||
      (int) rho.size () != nr_class * (nr_class - 1) / 2) return false;
  int sum = 0;
  for (int i = 0; i < nr_class; ++i) sum += nr_sv[i];
  if (sum != total_sv) return false;
  sv_coef.assign ((size_t) (nr_class - 1) * total_sv, 0.0);
  std::vector<std::vector<std::pair<int, double> > > rows (total_sv);
  int max_idx = min_dim;
  for (int s = 0; s < total_sv; ++s)
  {
    if (!std::getline (fs, line)) return false;
    std::istringstream is (line);
    for (int c = 0; c < nr_class - 1; ++c)
      if (!(is >> sv_coef[(size_t) c * total_sv + s])) return false;
    std::string tok;
    while (is >> tok)
    {
      const size_t colon = tok.find (':');
      if (colon == std::string::npos) return false;
      const int idx = std::atoi (tok.substr (0, colon).c_str ());
      const double val = std::strtod (tok.c_str () + colon + 1, 0);
      if (idx < 1) return false;
      rows[s].push_back (std::make_pair (idx, val));
      if (idx > max_idx) max_idx = idx;
    }
  }
  dim = max_idx;
  sv.assign ((size_t) total_sv * dim, 0.0);
  for (int s = 0; s < total_sv; ++s)
    for (size_t k = 0; k < rows[s].size (); ++k) sv[(size_t) s * dim + rows[s][k].first - 1] = rows[s][k].second;
  return true;
}

std::vector<std::vector<double> >
  SVMClassification::computeScaleParameters (const boost::shared_ptr<const InputType>& cloud, int startIdx, int nr_values)
{
  std::vector<std::vector<double> > res (2);
  res[0].assign (nr_values, DBL_MAX);
  res[1].assign (nr_values, -DBL_MAX);
  for (size_t cp = 0; cp < cloud->points.size (); cp++)
    for (int i = 0; i < nr_values; i++)
    {
      const double v = cloud->channels[startIdx + i].values[cp];
      if (res[0][i] > v) res[0][i] = v;
      else if (res[1][i] < v) res[1][i] = v;
    }
  return res;
}

bool
  SVMClassification::parseScaleParameterFile (const char *fileName, double &lower, double &upper, int nr_values,
                                              std::vector<std::vector<double> >& ranges, bool verbose)
{
  ranges.assign (2, std::vector<double> (nr_values, 0.0));
  std::ifstream fs;
  fs.open (fileName);
  if (!fs.is_open ())
  {
    if (verbose) ROS_ERROR ("Couldn't open %s for reading!", fileName);
    return false;
  }
  std::string mystring;
  fs >> mystring;
  if (mystring.substr (0, 1) != "x")
  {
    if (verbose) ROS_WARN ("X scaling not found in %s or unknown!", fileName);
    return false;
  }
  fs >> lower >> upper;
  int idx;
  float fmin, fmax;  // the reference reads the limits through floats (:168)
  while (fs >> idx >> fmin >> fmax)
    if (idx >= 1 && idx <= nr_values)
    {
      ranges[0][idx - 1] = fmin;
      ranges[1][idx - 1] = fmax;
    }
  fs.close ();
  return true;
}

void SVMClassification::init (ros::NodeHandle& nh)
{
  nh_ = nh;
}

void SVMClassification::pre ()
{
  nh_.param ("model_file_name", model_file_name_, model_file_name_);
  nh_.param ("scale_file_name", scale_file_name_, scale_file_name_);
  nh_.param ("scale_self", scale_self_, scale_self_);
  nh_.param ("scale_file", scale_file_, scale_file_);
}

void SVMClassification::post ()
{
}

std::vector<std::string> SVMClassification::requires ()
{
  std::vector<std::string> requires;
  requires.push_back ("f1");
  return requires;
}

std::vector<std::string> SVMClassification::provides ()
{
  std::vector<std::string> provides;
  provides.push_back ("point_class");
  return provides;
}

std::string SVMClassification::process (const boost::shared_ptr<const SVMClassification::InputType>& cloud)
{
  // Check if features exist and how many of them
  const int fIdx = getChannelIndex (cloud, "f1");
  if (fIdx == -1)
  {
    if (verbosity_level_ > -2) ROS_ERROR ("[SVMClassification] Provided point cloud does not have features computed. Use PFH or similar first!");
    output_valid_ = false;
    return std::string ("missing features");
  }
  int nr_values = 1;
  for (unsigned int d = fIdx + 1; d < cloud->channels.size (); d++)
  {
    char dim_name[16];
    std::snprintf (dim_name, sizeof (dim_name), "f%d", nr_values + 1);
    if (cloud->channels[d].name == dim_name) nr_values++;
  }
  const int plIdx = getChannelIndex (cloud, "point_label");

  SvmModelData model;
  if (!model.load (model_file_name_.c_str (), nr_values))
  {
    if (verbosity_level_ > -2) ROS_ERROR ("[SVMClassification] Couldn't load SVM model from %s", model_file_name_.c_str ());
    output_valid_ = false;
    return std::string ("incorrect model file");
  }

  double lower = 0, upper = 0;
  std::vector<std::vector<double> > value_ranges;
  if (scale_self_)
  {
    lower = -1;
    upper = +1;
    value_ranges = computeScaleParameters (cloud, fIdx, nr_values);
  }
  else if (scale_file_)
  {
    if (!parseScaleParameterFile (scale_file_name_.c_str (), lower, upper, nr_values, value_ranges, verbosity_level_ > -2))
    {
      if (verbosity_level_ > -2) ROS_ERROR ("[SVMClassification] Scaling requested from file %s but it is not possible!", scale_file_name_.c_str ());
      output_valid_ = false;
      return std::string ("incorrect scale parameter file");
    }
  }

  std::string err;
  cab_ctx* ctx = gpu_.get (err);
  if (!ctx) { output_valid_ = false; ROS_ERROR ("[SVMClassification] %s", err.c_str ()); return err; }

  ros::Time global_time = ros::Time::now ();
  cloud_svm_ = boost::shared_ptr<sensor_msgs::PointCloud> (new sensor_msgs::PointCloud ());
  cloud_svm_->header   = cloud->header;
  cloud_svm_->points   = cloud->points;
  cloud_svm_->channels = cloud->channels;
  const int pcIdx = (int) cloud_svm_->channels.size ();
  cloud_svm_->channels.resize (pcIdx + 1);
  cloud_svm_->channels[pcIdx].name = "point_class";
  const size_t n = cloud_svm_->points.size ();
  cloud_svm_->channels[pcIdx].values.resize (n, 0.0);

  // features point-major; dimensions the model knows but the cloud lacks are zero (absent sparse
  // entries in libsvm), dimensions beyond the model's are part of its (zero-padded) vectors
  const int dim = model.dim;
  std::vector<float> feat (n * (size_t) dim, 0.f);
  for (int i = 0; i < nr_values; i++)
    for (size_t cp = 0; cp < n; cp++) feat[cp * dim + i] = cloud->channels[fIdx + i].values[cp];
  int rc = cab_svm_set_model (ctx, dim, (int) model.labels.size (), model.total_sv, model.gamma, &model.labels[0], &model.nr_sv[0],
                              &model.rho[0], &model.sv_coef[0], &model.sv[0]);
  if (rc == CAB_OK)
  {
    if (!value_ranges.empty ())
    {
      // features beyond nr_values keep min == max == 0: scaleFeature returns 0 for them
      std::vector<double> fmin (dim, 0.0), fmax (dim, 0.0);
      for (int i = 0; i < nr_values; i++) { fmin[i] = value_ranges[0][i]; fmax[i] = value_ranges[1][i]; }
      rc = cab_svm_set_scaling (ctx, dim, lower, upper, &fmin[0], &fmax[0]);
    }
    else
      rc = cab_svm_set_scaling (ctx, dim, 0, 0, 0, 0);
  }
  if (rc == CAB_OK && n) rc = cab_svm_predict (ctx, &feat[0], (int64_t) n, dim, &cloud_svm_->channels[pcIdx].values[0], 0);
  if (rc != CAB_OK)
  {
    output_valid_ = false;
    err = std::string ("SVM classification failed: ") + cab_last_error (ctx);
    ROS_ERROR ("[SVMClassification] %s", err.c_str ());
    return err;
  }

  if (plIdx != -1)
  {
    int success = 0;
    for (size_t cp = 0; cp < n; cp++)
      if (cloud_svm_->channels[pcIdx].values[cp] == cloud_svm_->channels[plIdx].values[cp]) success++;
    if (verbosity_level_ > 0) ROS_INFO ("[SVMClassification] Accuracy: %d/%d (%g%%).", success, (int) n, success * 100.0 / n);
  }
  if (verbosity_level_ > 0) ROS_INFO ("[SVMClassification] SVM classification done in %g seconds.", (ros::Time::now () - global_time).toSec ());
  output_valid_ = true;
  return std::string ("ok");
}

boost::shared_ptr<const SVMClassification::OutputType> SVMClassification::output ()
  {return cloud_svm_;}

#ifdef CREATE_NODE
// the <algo>_node executable of the reference's CMakeLists.txt:42-57 (cloud_algos.h:106-117)
int main (int argc, char* argv[])
{
  return cloud_algos::standalone_node <cloud_algos::SVMClassification> (argc, argv);
}
#endif
