// nav_msgs/Path.h -- stand-in message struct.  TEST INFRASTRUCTURE ONLY; our own code.
#pragma once
#include <vector>
#include <geometry_msgs/PoseStamped.h>
namespace nav_msgs { struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; }; }
