// std_msgs/Header.h -- stand-in message struct (ROS is not installed here).  TEST INFRASTRUCTURE ONLY; our own code.
#pragma once
#include <string>
#include <ros/ros.h>
namespace std_msgs { struct Header { unsigned seq = 0; ros::Time stamp; std::string frame_id; }; }
