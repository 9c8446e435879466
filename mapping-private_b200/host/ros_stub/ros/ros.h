#pragma once
// Stand-in for ros/ros.h: a NodeHandle whose parameter server is an in-process map, publishers
// that count what they are given, and the logging macros.
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <boost/shared_ptr.hpp>

namespace ros {

inline int& log_level() {
  static int lvl = std::getenv("CLOUD_ALGOS_LOG") ? std::atoi(std::getenv("CLOUD_ALGOS_LOG")) : 0;
  return lvl;
}
#define ROS_INFO(...) do { if (::ros::log_level() > 1) { std::fprintf(stderr, "[INFO] " __VA_ARGS__); std::fputc('\n', stderr); } } while (0)
#define ROS_WARN(...) do { if (::ros::log_level() > 0) { std::fprintf(stderr, "[WARN] " __VA_ARGS__); std::fputc('\n', stderr); } } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, "[ERROR] " __VA_ARGS__); std::fputc('\n', stderr); } while (0)

struct Duration {
  double s = 0;
  double toSec() const { return s; }
};
struct Time {
  double t = 0;
  static Time now() {
    Time r;
    r.t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    return r;
  }
  Duration operator-(const Time& o) const { return Duration{t - o.t}; }
};

class Publisher {
 public:
  Publisher() : count_(std::make_shared<int>(0)) {}
  explicit Publisher(const std::string& topic) : topic_(topic), count_(std::make_shared<int>(0)) {}
  template <class M> void publish(const M&) const { ++*count_; }
  const std::string& getTopic() const { return topic_; }
  int getNumPublished() const { return *count_; }
 private:
  std::string topic_;
  std::shared_ptr<int> count_;
};

class Subscriber {};

class NodeHandle {
 public:
  explicit NodeHandle(const std::string& ns = "")
      : ns_(ns), params_(std::make_shared<std::map<std::string, double>>()), sparams_(std::make_shared<std::map<std::string, std::string>>()) {}
  // rosparam semantics: assign the stored value if present, else the default
  void param(const std::string& key, double& var, const double& def) const { var = has(key) ? (*params_)[key] : def; }
  void param(const std::string& key, int& var, const int& def) const { var = has(key) ? (int)(*params_)[key] : def; }
  void param(const std::string& key, bool& var, const bool& def) const { var = has(key) ? ((*params_)[key] != 0) : def; }
  void param(const std::string& key, std::string& var, const std::string& def) const {
    const std::string d = def;  // var and def may alias (nh.param("k", x_, x_))
    var = sparams_->count(key) ? (*sparams_)[key] : d;
  }
  void setParam(const std::string& key, double v) { (*params_)[key] = v; }
  void setParam(const std::string& key, const std::string& v) { (*sparams_)[key] = v; }
  void deleteParam(const std::string& key) { params_->erase(key); }
  bool hasParam(const std::string& key) const { return has(key); }
  template <class M> Publisher advertise(const std::string& topic, int /*queue*/) { return Publisher(topic); }
 private:
  bool has(const std::string& key) const { return params_->count(key) != 0; }
  std::string ns_;
  std::shared_ptr<std::map<std::string, double>> params_;  // shared between copies like a real handle
  std::shared_ptr<std::map<std::string, std::string>> sparams_;
};

}  // namespace ros
