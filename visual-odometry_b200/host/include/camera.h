// camera.h — drop-in for the reference's include/camera.h (class Camera, :11-63).
// Same constructor, same accessors, same projection semantics (integer z range, inclusive image
// bounds, invalid points marked (-1,-1)).  projectPoints() runs on the GPU
// (vo_project_points); projectPoint() is the one-point inline the reference also keeps inline.
#pragma once
#include "defs.h"

class Camera {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW

  Camera(int rows = 100, int cols = 100, int z_near = 0, int z_far = 10,
         const Eigen::Matrix3f& camera_matrix = Eigen::Matrix3f::Identity(),
         const Eigen::Isometry3f& world_in_camera_pose = Eigen::Isometry3f::Identity());

  // one point, on the host (reference camera.h:25-37): false when the point is outside the
  // depth range or the image
  inline bool projectPoint(Eigen::Vector2f& image_point, const Eigen::Vector3f& world_point) {
    const Eigen::Vector3f pc = _world_in_camera_pose * world_point;
    if (pc.z() > _z_far || pc.z() < _z_near) return false;
    const Eigen::Vector3f ph = _camera_matrix * pc;
    image_point = ph.head<2>() * (1. / ph.z());
    const bool inside_u = !(image_point.x() < 0 || image_point.x() > _cols - 1);
    const bool inside_v = !(image_point.y() < 0 || image_point.y() > _rows - 1);
    return inside_u && inside_v;
  }

  // a whole cloud, on the GPU (reference src/camera.cpp:16-37).  keep_indices: one output per
  // input with (-1,-1) for rejected points; otherwise the order-preserving compaction.
  // Returns the number of points inside the image.
  int projectPoints(Vector2fVector& image_points, const Vector3fVector& world_points,
                    bool keep_indices = false);

  inline const Eigen::Isometry3f& worldInCameraPose() const { return _world_in_camera_pose; }
  inline void setWorldInCameraPose(const Eigen::Isometry3f& pose) { _world_in_camera_pose = pose; }
  inline const Eigen::Matrix3f& cameraMatrix() const { return _camera_matrix; }
  inline const int rows() const { return _rows; }
  inline const int cols() const { return _cols; }
  // additions (the reference keeps these protected): needed to hand the camera to the device
  inline int zNear() const { return _z_near; }
  inline int zFar() const { return _z_far; }

 protected:
  int _rows, _cols;
  int _z_near, _z_far;
  Eigen::Matrix3f _camera_matrix;
  Eigen::Isometry3f _world_in_camera_pose;
};
