// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for <ros/ros.h> so that the reference's map_manager/src/Gridmap3D.cpp (plain grid arithmetic; its only uses of ROS are one
// ROS_ERROR in the never-called ESDF code and two unused ros::Time locals) can be compiled UNMODIFIED for oracle/_ref/libref_grid.so.
#pragma once
#include <cstdio>
#define ROS_ERROR(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_INFO(...) ((void)0)
namespace ros {
struct Time { static Time now() { return Time(); } double toSec() const { return 0.0; } };
}  // namespace ros
