// camera.cpp — host side of the drop-in Camera: construction and the GPU batch projection.
#include "camera.h"

#include "vo_b200_host.h"

Camera::Camera(int rows, int cols, int z_near, int z_far, const Eigen::Matrix3f& camera_matrix,
               const Eigen::Isometry3f& world_in_camera_pose)
    : _rows(rows),
      _cols(cols),
      _z_near(z_near),
      _z_far(z_far),
      _camera_matrix(camera_matrix),
      _world_in_camera_pose(world_in_camera_pose) {}

int Camera::projectPoints(Vector2fVector& image_points, const Vector3fVector& world_points,
                          bool keep_indices) {
  vo_camera cam;
  cam.rows = _rows;
  cam.cols = _cols;
  cam.z_near = _z_near;
  cam.z_far = _z_far;
  vo_b200::pack3(_camera_matrix, cam.K);
  vo_b200::pack_iso(_world_in_camera_pose, cam.T);
  image_points.resize(world_points.size());
  int64_t n_out = 0, n_inside = 0;
  if (!world_points.empty())
    vo_b200::check(vo_project_points(vo_b200::device(), &cam, world_points[0].data(),
                                     (int64_t)world_points.size(), keep_indices ? 1 : 0,
                                     image_points[0].data(), &n_out, &n_inside),
                   "vo_project_points");
  image_points.resize((size_t)n_out);
  return (int)n_inside;
}
