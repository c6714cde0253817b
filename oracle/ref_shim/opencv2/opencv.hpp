// opencv2/opencv.hpp -- the OpenCV names that estimator.h's headers declare with, and that Estimator::initialStructure
// (estimator.cpp:284-340, never called by the test driver) mentions.  Every function here aborts.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdlib>
#include <vector>
namespace cv {
class Mat {
 public:
  template <class T> Mat& operator<<(T) { return *this; }
  template <class T> Mat& operator,(T) { return *this; }
};
template <class T> class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int, int) {}
};
template <class T> struct Point_ { T x, y; Point_() : x(0), y(0) {} template <class A, class B> Point_(A a, B b) : x(T(a)), y(T(b)) {} };
template <class T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} template <class A, class B, class C> Point3_(A a, B b, C c) : x(T(a)), y(T(b)), z(T(c)) {} };
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
template <class A> void eigen2cv(const A&, Mat&) { std::abort(); }
template <class A> void cv2eigen(const Mat&, A&) { std::abort(); }
inline void Rodrigues(const Mat&, Mat&) { std::abort(); }
template <class A, class B> bool solvePnP(const A&, const B&, const Mat&, const Mat&, Mat&, Mat&, int) { std::abort(); }
}
