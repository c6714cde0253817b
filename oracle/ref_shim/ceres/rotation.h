// ceres/rotation.h -- declaration only: initial/initial_sfm.h names it inside a template that is never instantiated here.
#pragma once
namespace ceres { template <class T> void QuaternionRotatePoint(const T q[4], const T pt[3], T result[3]); }
