// ros/ros.h -- logging / assertion macros and the few ROS types the reference's headers mention (a NodeHandle that
// only stores string parameters, a Publisher nobody listens to).  ROS is not installed in this image.
// TEST INFRASTRUCTURE ONLY; our own code.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <map>
#include <sstream>
#include <string>
#define ROS_INFO(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_WARN(...) do { std::fprintf(stderr, "[ref WARN] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, "[ref ERROR] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_DEBUG_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_ERROR_STREAM(x) ((void)0)
#define ROS_BREAK() std::abort()
#define ROS_ASSERT(c) do { if (!(c)) { std::fprintf(stderr, "ROS_ASSERT failed: %s\n", #c); std::abort(); } } while (0)
namespace ros {
struct Time {
  unsigned sec = 0, nsec = 0;
  Time() {}
  explicit Time(double t) : sec(unsigned(t)), nsec(unsigned((t - unsigned(t)) * 1e9)) {}
  double toSec() const { return sec + 1e-9 * nsec; }
};
class Publisher {
 public:
  int getNumSubscribers() const { return 0; }
  template <class M> void publish(const M&) const {}
};
class NodeHandle {
  std::map<std::string, std::string> params_;
 public:
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  void setParam(const std::string& k, const std::string& v) { params_[k] = v; }
  bool getParam(const std::string& k, std::string& v) const { auto it = params_.find(k); if (it == params_.end()) return false; v = it->second; return true; }
  template <class M> Publisher advertise(const std::string&, int) { return Publisher(); }
};
}  // namespace ros
