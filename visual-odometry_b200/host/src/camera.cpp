// camera.cpp — host side of the drop-in Camera: construction and the GPU batch projection.
#include "camera.h"

#include "vo_b200_host.h"

Camera::Camera(int height, int width, int near_plane, int far_plane, const Eigen::Matrix3f& K,
               const Eigen::Isometry3f& pose)
    : _pose(pose), _K(K), _height(height), _width(width), _near(near_plane), _far(far_plane) {}

// Camera::projectPoints of the reference (src/camera.cpp:16-37), executed by project_points_kernel
int Camera::projectPoints(Vector2fVector& pixels, const Vector3fVector& world_points, bool keep_indices) {
  const int64_t n = (int64_t)world_points.size();
  pixels.resize(world_points.size());
  if (n == 0) return 0;
  vo_camera description;
  description.rows = _height;
  description.cols = _width;
  description.z_near = _near;
  description.z_far = _far;
  vo_b200::pack3(_K, description.K);
  vo_b200::pack_iso(_pose, description.T);
  int64_t written = 0, accepted = 0;
  vo_b200::check(vo_project_points(vo_b200::device(), &description, world_points[0].data(), n,
                                   keep_indices ? 1 : 0, pixels[0].data(), &written, &accepted),
                 "vo_project_points");
  pixels.resize((size_t)written);
  return (int)accepted;
}
