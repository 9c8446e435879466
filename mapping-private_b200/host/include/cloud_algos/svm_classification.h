// svm_classification.h -- cloud_algos::SVMClassification on the B200.
// Same public surface as cloud_algos/include/cloud_algos/svm_classification.h of the reference
// (:27-70 options, topics, defaults; :64-86 scaleFeature; :88-127 computeScaleParameters;
// :129-185 parseScaleParameterFile).  The reference links libsvm (svm_load_model / svm_predict);
// here the model file is parsed on the host and prediction runs on the GPU through
// cab_svm_set_model / cab_svm_set_scaling / cab_svm_predict (include/cloud_algos_b200.h).
#ifndef CLOUD_ALGOS_SVM_CLASSIFICATION_H
#define CLOUD_ALGOS_SVM_CLASSIFICATION_H
#include <cloud_algos/cloud_algos.h>

#include <float.h>
#include <string>
#include <vector>

namespace cloud_algos
{

// A libsvm C-SVC / RBF model as plain arrays (what svm_load_model holds for the svm/*.model files).
struct SvmModelData
{
  double gamma;
  std::vector<int> labels, nr_sv;      // nr_class each
  std::vector<double> rho;             // nr_class * (nr_class - 1) / 2
  std::vector<double> sv_coef;         // (nr_class - 1) x total_sv
  std::vector<double> sv;              // total_sv x dim, dense
  int dim, total_sv;
  SvmModelData () : gamma (0), dim (0), total_sv (0) {}
  // false if the file cannot be read or is not a c_svc / rbf model; min_dim pads the vectors
  bool load (const char* file_name, int min_dim);
};

class SVMClassification : public CloudAlgo
{
 public:
  typedef sensor_msgs::PointCloud OutputType;
  typedef sensor_msgs::PointCloud InputType;

  // Options
  std::string model_file_name_; // filename where the model should be loaded from
  std::string scale_file_name_; // filename where the scale parameters should be loaded from
  bool scale_self_;             // scale every feature with its own minimum to -1 and maximum to 1
  bool scale_file_;             // if scale_self_ is off: use the ranges from scale_file_name_

  static std::string default_input_topic () {return std::string ("cloud_pcd");}
  static std::string default_output_topic () {return std::string ("cloud_svm");}
  static std::string default_node_name () {return std::string ("svm_classification_node");}

  void init (ros::NodeHandle& nh);
  void pre ();
  void post ();
  std::vector<std::string> requires ();
  std::vector<std::string> provides ();
  std::string process (const boost::shared_ptr<const InputType>&);
  boost::shared_ptr<const OutputType> output ();

  SVMClassification () : CloudAlgo ()
  {
    model_file_name_ = std::string ("svm/fpfh.model");
    scale_file_name_ = std::string ("svm/teapot_smooth_fpfh.scp");
    scale_self_ = false;
    scale_file_ = true; // gets considered only if scale_self is false
  }

  // min / max of every feature channel; result[0] = minima, result[1] = maxima (2 x nr_values).
  // Keeps the reference's update rule (a value that lowers the minimum is not tested against the
  // maximum, svm_classification.h:112-117).
  static std::vector<std::vector<double> >
    computeScaleParameters (const boost::shared_ptr<const InputType>& cloud, int startIdx, int nr_values);

  // false on failure (file missing, no "x" section); ranges[0] / ranges[1] as above
  static bool
    parseScaleParameterFile (const char *fileName, double &lower, double &upper, int nr_values,
                             std::vector<std::vector<double> >& ranges, bool verbose = true);

  ros::Publisher createPublisher (ros::NodeHandle& nh)
  {
    ros::Publisher p = nh.advertise<OutputType> (default_output_topic (), 5);
    return p;
  }
 private:
  ros::NodeHandle nh_;
  boost::shared_ptr<sensor_msgs::PointCloud> cloud_svm_;
  GpuContext gpu_;
};

}
#endif
