#pragma once
// Stand-in for ros/ros.h: a NodeHandle whose parameter server is an in-process map, publishers
// that count what they are given, an in-process topic table (subscribe / inject / spin) and the logging macros.
#include <chrono>
#include <deque>
#include <functional>
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

// In-process stand-in of the master: topics that have a subscriber and the messages waiting for them.  A test (or
// the node mains built with -DCREATE_NODE) injects a message, ros::spin() delivers what is queued and returns.
namespace master {
struct TopicInfo {
  std::string name, datatype;
};
struct Table {
  std::map<std::string, std::function<void(const boost::shared_ptr<const void>&)>> subscribers;
  std::deque<std::pair<std::string, boost::shared_ptr<const void>>> queue;
  std::vector<TopicInfo> advertised;
};
inline Table& table() {
  static Table t;
  return t;
}
inline bool getTopics(std::vector<TopicInfo>& out) {
  out = table().advertised;
  return true;
}
}  // namespace master

namespace this_node {
inline std::string& name_ref() {
  static std::string n = "node";
  return n;
}
inline const std::string& getName() { return name_ref(); }
}  // namespace this_node

inline void init(int&, char**, const std::string& name) { this_node::name_ref() = name; }
template <class M>
inline void inject(const std::string& topic, const boost::shared_ptr<const M>& msg) {
  master::table().advertised.push_back(master::TopicInfo{topic, ""});
  master::table().queue.emplace_back(topic, boost::shared_ptr<const void>(msg));
}
// delivers the queued messages to their subscribers, then returns (a real ros::spin() blocks until shutdown)
inline void spin() {
  master::Table& t = master::table();
  while (!t.queue.empty()) {
    auto item = t.queue.front();
    t.queue.pop_front();
    auto it = t.subscribers.find(item.first);
    if (it != t.subscribers.end()) it->second(item.second);
  }
}

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
  template <class M, class T>
  Subscriber subscribe(const std::string& topic, int /*queue*/, void (T::*cb)(const boost::shared_ptr<const M>&), T* obj) {
    master::table().subscribers[topic] = [cb, obj](const boost::shared_ptr<const void>& m) {
      (obj->*cb)(boost::static_pointer_cast<const M>(m));
    };
    return Subscriber();
  }
 private:
  bool has(const std::string& key) const { return params_->count(key) != 0; }
  std::string ns_;
  std::shared_ptr<std::map<std::string, double>> params_;  // shared between copies like a real handle
  std::shared_ptr<std::map<std::string, std::string>> sparams_;
};

}  // namespace ros
