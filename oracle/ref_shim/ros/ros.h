// ros/ros.h -- logging / assertion macros and the one type name the reference's factor headers mention.  ROS is not
// installed in this image.  TEST INFRASTRUCTURE ONLY; our own code.
#pragma once
#include <cstdio>
#include <cstdlib>
#define ROS_INFO(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_WARN(...) do { std::fprintf(stderr, "[ref WARN] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, "[ref ERROR] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_DEBUG_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_BREAK() std::abort()
#define ROS_ASSERT(c) do { if (!(c)) { std::fprintf(stderr, "ROS_ASSERT failed: %s\n", #c); std::abort(); } } while (0)
namespace ros { class NodeHandle; }
