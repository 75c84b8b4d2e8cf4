// camera.h — drop-in for the reference's include/camera.h (class Camera, :11-63).
// The public surface is the reference's (constructor, projectPoint, projectPoints, the pose and
// matrix accessors, rows/cols); projectPoints() runs on the GPU through vo_project_points, the
// single-point projectPoint() stays a host inline.  Semantics kept: integer depth range, inclusive
// image bounds, rejected points marked (-1,-1) by the batch call.
#pragma once
#include "defs.h"

class Camera {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW

  Camera(int height = 100, int width = 100, int near_plane = 0, int far_plane = 10,
         const Eigen::Matrix3f& K = Eigen::Matrix3f::Identity(),
         const Eigen::Isometry3f& pose = Eigen::Isometry3f::Identity());

  // state
  const Eigen::Matrix3f& cameraMatrix() const { return _K; }
  const Eigen::Isometry3f& worldInCameraPose() const { return _pose; }
  void setWorldInCameraPose(const Eigen::Isometry3f& pose) { _pose = pose; }
  const int rows() const { return _height; }
  const int cols() const { return _width; }
  // additions: the depth range, needed to describe the camera to the device
  int zNear() const { return _near; }
  int zFar() const { return _far; }

  // one world point -> pixel, on the host; false when it is outside the depth range or the image
  bool projectPoint(Eigen::Vector2f& pixel, const Eigen::Vector3f& world_point);

  // a whole cloud, on the GPU.  keep_indices: one output per input, (-1,-1) where rejected;
  // otherwise only the accepted pixels, in input order.  Returns how many were accepted.
  int projectPoints(Vector2fVector& pixels, const Vector3fVector& world_points,
                    bool keep_indices = false);

 protected:
  Eigen::Isometry3f _pose;  // world in camera
  Eigen::Matrix3f _K;
  int _height, _width;      // image size in pixels
  int _near, _far;          // depth range (integers, as in the reference)
};

inline bool Camera::projectPoint(Eigen::Vector2f& pixel, const Eigen::Vector3f& world_point) {
  const Eigen::Vector3f in_camera = _pose * world_point;
  const float depth = in_camera.z();
  if (depth > _far || depth < _near) return false;
  const Eigen::Vector3f h = _K * in_camera;
  pixel = h.head<2>() * (1. / h.z());  // double reciprocal, then demoted: the reference's rounding
  const bool u_ok = !(pixel.x() < 0 || pixel.x() > _width - 1);
  const bool v_ok = !(pixel.y() < 0 || pixel.y() > _height - 1);
  return u_ok && v_ok;
}
