#pragma once
// Stand-in for pluginlib: PLUGINLIB_DECLARE_CLASS registers a factory under "<pkg>/<class_name>",
// pluginlib::ClassLoader<Base> looks it up (the calls table_memory_grsd.cpp:683-705 makes).
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace pluginlib {

class PluginlibException : public std::runtime_error {
 public:
  explicit PluginlibException(const std::string& m) : std::runtime_error(m) {}
};

template <class Base>
struct Registry {
  static std::map<std::string, std::function<Base*()>>& map() {
    static std::map<std::string, std::function<Base*()>> m;
    return m;
  }
};

template <class Base>
class ClassLoader {
 public:
  ClassLoader(const std::string& package, const std::string& base_class) : package_(package), base_(base_class) {}
  void loadLibraryForClass(const std::string& lookup_name) {
    if (!Registry<Base>::map().count(lookup_name))
      throw PluginlibException("no class " + lookup_name + " with base " + base_ + " in package " + package_);
  }
  Base* createClassInstance(const std::string& lookup_name, bool /*auto_load*/ = true) {
    auto it = Registry<Base>::map().find(lookup_name);
    if (it == Registry<Base>::map().end()) throw PluginlibException("no class " + lookup_name);
    return it->second();
  }
  std::vector<std::string> getDeclaredClasses() const {
    std::vector<std::string> v;
    for (auto& kv : Registry<Base>::map()) v.push_back(kv.first);
    return v;
  }
 private:
  std::string package_, base_;
};

template <class Base, class Derived>
struct Registrar {
  explicit Registrar(const char* name) { Registry<Base>::map()[name] = []() -> Base* { return new Derived(); }; }
};

}  // namespace pluginlib

#define PLUGINLIB_CONCAT2(a, b) a##b
#define PLUGINLIB_CONCAT(a, b) PLUGINLIB_CONCAT2(a, b)
#define PLUGINLIB_DECLARE_CLASS(pkg, class_name, class_type, base_class_type) \
  static ::pluginlib::Registrar<base_class_type, class_type> PLUGINLIB_CONCAT(pluginlib_registrar_, __LINE__)(#pkg "/" #class_name)
