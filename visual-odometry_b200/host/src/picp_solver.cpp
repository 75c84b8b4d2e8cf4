// picp_solver.cpp — host side of the drop-in PICPSolver: marshals the reference's calls onto a
// vo_picp handle.  All arithmetic happens in visual-odometry_b200/csrc/picp.cu.
#include "picp_solver.h"

#include <cstring>

#include "vo_b200_host.h"

using vo_b200::check;

PICPSolver::PICPSolver()
    : _handle(nullptr),
      _kernel_thereshold(1000.f),  // reference picp_solver.cpp:13
      _damping(1.f),               // :10
      _min_num_inliers(0),         // :11
      _dirty(false),
      _pending_rounds(0),
      _pending_keep(false),
      _chi_inliers(0.f),
      _chi_outliers(0.f),
      _num_inliers(0) {
  _H.setZero();
  _b.setZero();
  check(vo_picp_create(&_handle, vo_b200::device()), "vo_picp_create");
}

PICPSolver::~PICPSolver() { vo_picp_destroy(_handle); }

void PICPSolver::setKernelThreshold(float kernel_threshold) {
  _kernel_thereshold = kernel_threshold;
  check(vo_picp_set_params(_handle, _kernel_thereshold, _damping, _min_num_inliers),
        "vo_picp_set_params");
}

void PICPSolver::init(const Camera& camera, const Vector3fVector& world_points,
                      const Vector2fVector& image_points) {
  _pending_rounds = 0;  // rounds queued on the previous problem were never observed
  _camera = camera;
  vo_camera cam;
  cam.rows = camera.rows();
  cam.cols = camera.cols();
  cam.z_near = camera.zNear();
  cam.z_far = camera.zFar();
  vo_b200::pack3(camera.cameraMatrix(), cam.K);
  vo_b200::pack_iso(camera.worldInCameraPose(), cam.T);
  check(vo_picp_set_params(_handle, _kernel_thereshold, _damping, _min_num_inliers),
        "vo_picp_set_params");
  check(vo_picp_init(_handle, &cam, world_points.empty() ? nullptr : world_points[0].data(),
                     (int64_t)world_points.size(),
                     image_points.empty() ? nullptr : image_points[0].data(),
                     (int64_t)image_points.size()),
        "vo_picp_init");
  _pairs_cache.clear();
  _pairs_cache.push_back(-1);  // never equal to a real upload: forces the first one
  _dirty = false;
  _chi_inliers = _chi_outliers = 0.f;
  _num_inliers = 0;
}

// The reference receives the correspondence vector on every oneRound() call (typically the same
// one a hundred times).  Re-uploading it each round would cost more than the round itself, so
// the device copy is refreshed only when the contents differ from the last upload.
bool PICPSolver::upload(const IntPairVector& correspondences) {
  const size_t n = correspondences.size();
  const int* flat = n ? &correspondences[0].first : nullptr;
  if (_pairs_cache.size() == 2 * n && (n == 0 || std::memcmp(_pairs_cache.data(), flat, 8 * n) == 0))
    return false;
  flush();  // queued rounds belong to the previous correspondences
  check(vo_picp_set_correspondences(_handle, flat, (int64_t)n), "vo_picp_set_correspondences");
  _pairs_cache.assign(flat, flat + 2 * n);
  return true;
}

void PICPSolver::flush() const {
  if (_pending_rounds == 0) return;
  const int rounds = _pending_rounds;
  _pending_rounds = 0;
  check(vo_picp_compute(_handle, _pending_keep ? 1 : 0, rounds), "vo_picp_compute");
  _dirty = true;
}

bool PICPSolver::compute(const IntPairVector& correspondences, bool keep_outliers, int rounds) {
  flush();
  upload(correspondences);
  check(vo_picp_compute(_handle, keep_outliers ? 1 : 0, rounds), "vo_picp_compute");
  _dirty = true;
  // the reference returns false only when fewer than _min_num_inliers points are inliers; that
  // member is 0 and has no setter (picp_solver.cpp:11,103-107), so the round always succeeds and
  // there is nothing to wait for
  if (_min_num_inliers <= 0) return true;
  refresh();
  return _num_inliers >= _min_num_inliers;
}

bool PICPSolver::oneRound(const IntPairVector& correspondences, bool keep_outliers) {
  // with a minimum inlier count the return value depends on the round itself: nothing to defer
  if (_min_num_inliers > 0) return compute(correspondences, keep_outliers, 1);
  if (_pending_rounds > 0 && keep_outliers != _pending_keep) flush();
  upload(correspondences);  // flushes first when the correspondences changed
  _pending_keep = keep_outliers;
  ++_pending_rounds;
  return true;  // picp_solver.cpp:103-107 can only fail below _min_num_inliers (0, no setter)
}

void PICPSolver::refresh() const {
  flush();
  if (!_dirty) return;
  vo_picp_state st;
  check(vo_picp_get_state(_handle, &st), "vo_picp_get_state");
  Eigen::Isometry3f pose = Eigen::Isometry3f::Identity();
  vo_b200::unpack_iso(st.T, pose);
  _camera.setWorldInCameraPose(pose);
  _chi_inliers = st.chi_inliers;
  _chi_outliers = st.chi_outliers;
  _num_inliers = st.num_inliers;
  for (int j = 0; j < 6; ++j)
    for (int i = 0; i < 6; ++i) _H(i, j) = st.H[j * 6 + i];
  for (int i = 0; i < 6; ++i) _b(i) = st.b[i];
  _dirty = false;
}

const Camera& PICPSolver::camera() const {
  refresh();
  return _camera;
}
const float PICPSolver::chiInliers() const {
  refresh();
  return _chi_inliers;
}
const float PICPSolver::chiOutliers() const {
  refresh();
  return _chi_outliers;
}
const int PICPSolver::numInliers() const {
  refresh();
  return _num_inliers;
}
