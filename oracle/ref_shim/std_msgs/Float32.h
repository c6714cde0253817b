// std_msgs/Float32.h -- stand-in message struct.  TEST INFRASTRUCTURE ONLY; our own code.
#pragma once
namespace std_msgs { struct Float32 { float data = 0; }; }
