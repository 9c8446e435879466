// register_plugins.cpp -- pluginlib registration (reference: cloud_algos/src/register_plugins.cpp:15-24).
// The hot-path plugins and the SVM classifier that consumes GRSD are registered under the lookup names callers use
// ("cloud_algos/LocalRadiusEstimation", "cloud_algos/GlobalRSD": table_memory_grsd.cpp:288,852).
#include <pluginlib/class_list_macros.h>
#include <cloud_algos/cloud_algos.h>
#include <cloud_algos/normal_estimation.h>
#include <cloud_algos/radius_estimation.h>
#include <cloud_algos/global_rsd.h>
#include <cloud_algos/svm_classification.h>
#include <cloud_algos/noise_removal.h>
#include <cloud_algos/pfh.h>

using namespace cloud_algos;

PLUGINLIB_DECLARE_CLASS(cloud_algos, NormalEstimation, cloud_algos::NormalEstimation, cloud_algos::CloudAlgo);
PLUGINLIB_DECLARE_CLASS(cloud_algos, LocalRadiusEstimation, cloud_algos::LocalRadiusEstimation, cloud_algos::CloudAlgo);
PLUGINLIB_DECLARE_CLASS(cloud_algos, GlobalRSD, cloud_algos::GlobalRSD, cloud_algos::CloudAlgo);
PLUGINLIB_DECLARE_CLASS(cloud_algos, SVMClassification, cloud_algos::SVMClassification, cloud_algos::CloudAlgo);
PLUGINLIB_DECLARE_CLASS(cloud_algos, StatisticalNoiseRemoval, cloud_algos::StatisticalNoiseRemoval, cloud_algos::CloudAlgo);
PLUGINLIB_DECLARE_CLASS(cloud_algos, PointFeatureHistogram, cloud_algos::PointFeatureHistogram, cloud_algos::CloudAlgo);
