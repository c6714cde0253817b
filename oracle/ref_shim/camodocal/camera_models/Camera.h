// camodocal/camera_models/Camera.h -- stand-in for the camera interface FeatureSelector uses (spaceToPlane, image size).
// The reference vendors camodocal (camera_model/), but its sources need OpenCV, which is not installed here; the one
// model the EuRoC configuration uses is restated below from camera_model/src/camera_models/PinholeCamera.cc:520-542
// (spaceToPlane) and :646-662 (radial-tangential distortion).  TEST INFRASTRUCTURE ONLY; our own code.
#pragma once
#include <memory>
#include <string>
#include <eigen3/Eigen/Dense>
namespace camodocal {
class Camera {
 public:
  virtual ~Camera() {}
  virtual int imageWidth(void) const = 0;
  virtual int imageHeight(void) const = 0;
  virtual void spaceToPlane(const Eigen::Vector3d& P, Eigen::Vector2d& p) const = 0;
};
typedef std::shared_ptr<Camera> CameraPtr;
struct PinholeParams { double fx, fy, cx, cy, k1, k2, p1, p2; int width, height; };
class PinholeCameraShim : public Camera {
  PinholeParams c_;
 public:
  explicit PinholeCameraShim(const PinholeParams& c) : c_(c) {}
  const PinholeParams& params() const { return c_; }   // camodocal's PinholeCamera::getParameters()
  int imageWidth(void) const override { return c_.width; }
  int imageHeight(void) const override { return c_.height; }
  void spaceToPlane(const Eigen::Vector3d& P, Eigen::Vector2d& p) const override {
    const double mx = P(0) / P(2), my = P(1) / P(2);
    const double mx2 = mx * mx, my2 = my * my, mxy = mx * my, rho2 = mx2 + my2, rad = c_.k1 * rho2 + c_.k2 * rho2 * rho2;
    const double dx = mx * rad + 2.0 * c_.p1 * mxy + c_.p2 * (rho2 + 2.0 * mx2);
    const double dy = my * rad + 2.0 * c_.p2 * mxy + c_.p1 * (rho2 + 2.0 * my2);
    p(0) = c_.fx * (mx + dx) + c_.cx; p(1) = c_.fy * (my + dy) + c_.cy;
  }
};
}  // namespace camodocal
