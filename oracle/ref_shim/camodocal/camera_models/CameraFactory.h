// camodocal/camera_models/CameraFactory.h -- stand-in: the "calibration file" is the parameter block the test driver
// registered beforehand (there is no YAML reader here).  TEST INFRASTRUCTURE ONLY; our own code.
#pragma once
#include "camodocal/camera_models/Camera.h"
namespace camodocal {
class CameraFactory {
 public:
  static PinholeParams& registered() { static PinholeParams p{}; return p; }
  static std::shared_ptr<CameraFactory> instance(void) { static std::shared_ptr<CameraFactory> f(new CameraFactory()); return f; }
  CameraPtr generateCameraFromYamlFile(const std::string&) { return CameraPtr(new PinholeCameraShim(registered())); }
};
}  // namespace camodocal
