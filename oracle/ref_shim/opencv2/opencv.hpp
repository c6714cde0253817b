// opencv2/opencv.hpp -- the type names that appear in DECLARATIONS of headers estimator.h pulls in
// (initial/solve_5pts.h, initial/initial_ex_rotation.h).  Nothing here is ever called.  TEST INFRASTRUCTURE ONLY.
#pragma once
namespace cv {
class Mat {};
template <class T> class Mat_ : public Mat {};
template <class T> struct Point_ { T x, y; };
template <class T> struct Point3_ { T x, y, z; };
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
}
