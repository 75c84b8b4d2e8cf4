// vo_sequence.cpp — the vo_complete pipeline on an in-memory SYNTHETIC sequence (BASELINE config 5).
//
// One source, two builds:
//   * against this repository's drop-in headers + libvo_b200.so  -> host/bin/vo_sequence (B200)
//   * against the reference's own headers and sources             -> oracle/_ref/bin/vo_sequence (CPU)
// It uses only declarations both header sets share (TreeNode_, PICPSolver, triangulate_points,
// estimate_transform, PointCloudVector, Camera), so the second build times the REFERENCE's
// implementation of every stage on exactly the same frames.  The per-frame loop follows the
// reference's main (src/apps/vo_complete.cpp:116-178): appearance association of consecutive
// frames, join with the previous triangulation, 100 PICP rounds, re-triangulation, map update.
//
// Synthetic data: landmarks uniform in the 20 m x 20 m x 2 m volume of
// the bundled world.dat with U(-1,1)^10 appearances; a planar robot trajectory with 0.2 m steps
// and small turns that stays inside the volume; the bundled camera (camera.dat: K, cam_transform,
// z in (0,5), 640x480); a frame lists the visible landmarks in ascending id, noise-free, with the
// landmark's appearance copied verbatim.  Frames are generated on the fly, outside the timed region.
//
//   vo_sequence <n_landmarks> <n_frames> [seed=1000] [rounds=100] [pose_dump_file]
//   env: VO_SEQ_MODE=classes (GPU build: use the drop-in classes instead of the resident pipeline),
//        VO_SEQ_MAPDUMP=<file> (write the final map), VO_SEQ_LOG=1, VO_SEQ_THETA0=<rad>
// prints one JSON line (frames/s over the frame loop, per-stage milliseconds, accuracy vs GT).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "utils.h"
#include "camera.h"
#include "picp_solver.h"
#include "eigen_kdtree.h"
#include "epipolar_utils.h"

namespace {

using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

// untimed frames at the start of the loop (both builds): a tenth of the sequence, at most 20
int warmup_frames(int n_frames) { return std::min(20, (n_frames - 2) / 10); }

struct World {
  Vector3fVector points;
  Vector10fVector appearances;
};

World make_world(int n, unsigned seed) {
  std::mt19937 rng(seed);
  std::uniform_real_distribution<float> xy(-10.f, 10.f), z(0.f, 2.f), app(-1.f, 1.f);
  World w;
  w.points.resize(n);
  w.appearances.resize(n);
  for (int i = 0; i < n; ++i) {
    w.points[i] = Eigen::Vector3f(xy(rng), xy(rng), z(rng));
    for (int d = 0; d < 10; ++d) w.appearances[i](d) = app(rng);
  }
  return w;
}

Eigen::Isometry3f planar_pose(float x, float y, float theta) {
  Eigen::Isometry3f X = Eigen::Isometry3f::Identity();
  Eigen::Matrix3f R;
  const float c = std::cos(theta), s = std::sin(theta);
  R << c, -s, 0.f, s, c, 0.f, 0.f, 0.f, 1.f;
  X.linear() = R;
  X.translation() = Eigen::Vector3f(x, y, 0.f);
  return X;
}

// cam_transform of the bundled camera.dat: the camera looks along the robot's x axis
Eigen::Isometry3f camera_in_robot() {
  Eigen::Isometry3f X = Eigen::Isometry3f::Identity();
  Eigen::Matrix3f R;
  R << 0.f, 0.f, 1.f, -1.f, 0.f, 0.f, 0.f, -1.f, 0.f;
  X.linear() = R;
  X.translation() = Eigen::Vector3f(0.2f, 0.f, 0.f);
  return X;
}

// The bundled robot (trajectory.dat) either drives straight in 0.2 m steps or turns on the spot
// by 0.2 rad per frame; this one does the same: straight legs across the landmark volume and,
// when it gets close to the border heading outwards, an in-place turn back towards the centre.
struct Robot {
  float x = 0.f, y = 0.f, theta = 0.f;
  std::mt19937 rng;
  float target = 0.f;
  bool turning = false;
  explicit Robot(unsigned seed) : rng(seed) {
    if (const char* e = std::getenv("VO_SEQ_THETA0")) theta = (float)std::atof(e);
  }
  static float wrap(float a) {
    while (a > 3.14159265f) a -= 6.2831853f;
    while (a < -3.14159265f) a += 6.2831853f;
    return a;
  }
  void step() {
    if (!turning && std::hypot(x, y) > 6.f && (x * std::cos(theta) + y * std::sin(theta)) > 0.f) {
      std::uniform_real_distribution<float> jitter(-0.6f, 0.6f);
      target = wrap(std::atan2(-y, -x) + jitter(rng));  // somewhere across the volume
      turning = true;
    }
    if (turning) {
      const float d = wrap(target - theta);
      const float turn = std::max(-0.2f, std::min(0.2f, d));
      theta = wrap(theta + turn);
      if (std::fabs(d) <= 0.2f) turning = false;
      return;  // turn on the spot
    }
    x += 0.2f * std::cos(theta);
    y += 0.2f * std::sin(theta);
  }
};

struct Frame {
  PointCloudVector<2> pc;
  Eigen::Isometry3f world_in_camera;
};

// Measurement synthesis: ONE batch call of Camera::projectPoints over all the
// landmarks (src/camera.cpp:16-37, keep_indices = true: rejected points come back as (-1,-1)) instead
// of a host loop over projectPoint — in the drop-in build that is project_points_kernel on the GPU,
// in the reference build the reference's own loop; both produce the same bits.  Visible landmarks
// are listed in ascending id with their appearance copied verbatim, as in the bundled meas-*.dat.
Frame observe(const World& w, Camera& cam, const Robot& r) {
  Frame f;
  f.world_in_camera = (planar_pose(r.x, r.y, r.theta) * camera_in_robot()).inverse();
  cam.setWorldInCameraPose(f.world_in_camera);
  Vector2fVector uv;
  cam.projectPoints(uv, w.points, true);
  for (size_t i = 0; i < w.points.size(); ++i)
    if (uv[i].x() >= 0.f) f.pc.push_back(PointCloud<2>(uv[i], w.appearances[i]));
  return f;
}

// appearance association of two frames -> (index in frame 1, index in frame 2), in query order.
// Same contract as the reference's compute_correspondences_images (vo_complete.cpp:12-48): the
// search structure is built over the larger set, the smaller set queries it with radius 0.1.
IntPairVector associate(const Vector10fVector& app1, const Vector10fVector& app2) {
  using Container = Vector11fVector;
  using Tree = TreeNode_<Container::iterator>;
  const bool first_is_map = app1.size() >= app2.size();
  const Vector10fVector& map_app = first_is_map ? app1 : app2;
  const Vector10fVector& query_app = first_is_map ? app2 : app1;
  Container map_rows(map_app.size()), queries(query_app.size());
  for (size_t i = 0; i < map_rows.size(); ++i) map_rows[i] << float(i), map_app[i];
  for (size_t i = 0; i < queries.size(); ++i) queries[i] << float(i), query_app[i];
  IntPairVector out;
  out.reserve(queries.size());
  if (map_rows.empty() || queries.empty()) return out;
  Tree tree(map_rows.begin(), map_rows.end(), 10);
#ifdef VO_B200_DROPIN
  std::vector<int> hit;
  tree.bestMatchFullBatch(queries.begin(), queries.end(), 0.1f, hit);  // one launch for the frame
  for (size_t q = 0; q < queries.size(); ++q) {
    if (hit[q] < 0) continue;
    const int m = int(map_rows[hit[q]](0));
    out.push_back(first_is_map ? IntPair(m, int(q)) : IntPair(int(q), m));
  }
#else
  for (size_t q = 0; q < queries.size(); ++q) {
    Vector11f* best = tree.bestMatchFull(queries[q], 0.1f);
    if (!best) continue;
    const int m = int((*best)(0));
    out.push_back(first_is_map ? IntPair(m, int(q)) : IntPair(int(q), m));
  }
#endif
  return out;
}

// (ref,cur) x (ref,world) -> (cur,world): the FIRST world entry of each reference index wins
// (the contract of extract_correspondences_world, vo_complete.cpp:51-66)
IntPairVector join_with_world(const IntPairVector& ref_cur, const IntPairVector& ref_world) {
  int max_ref = -1;
  for (const IntPair& rw : ref_world) max_ref = std::max(max_ref, rw.first);
  std::vector<int> first_world((size_t)(max_ref + 1), -1);
  for (const IntPair& rw : ref_world)
    if (first_world[rw.first] < 0) first_world[rw.first] = rw.second;  // keeps the first
  IntPairVector out;
  out.reserve(ref_cur.size());
  for (const IntPair& rc : ref_cur)
    if (rc.first >= 0 && rc.first <= max_ref && first_world[rc.first] >= 0)
      out.push_back(IntPair(rc.second, first_world[rc.first]));
  return out;
}

// rotation angle of R (small angles: the antisymmetric part is sin(angle) * axis; acos of the
// trace would lose everything below 3e-4 rad in FP32)
float rotation_angle(const Eigen::Matrix3f& R) {
  const float a = R(2, 1) - R(1, 2), b = R(0, 2) - R(2, 0), c = R(1, 0) - R(0, 1);
  const float s = 0.5f * std::sqrt(a * a + b * b + c * c);
  return std::asin(std::min(1.f, s));
}


// frames (in order) before the estimated step length first falls below half of its initial value:
// the reference's monocular scale, carried from frame to frame by a two-view triangulation, decays
// on long sequences in BOTH builds and then collapses to a rotation-only fixed point
int scale_alive_frames(const std::vector<double>& ratio_in_order) {
  if (ratio_in_order.empty()) return 0;
  const double r0 = ratio_in_order.front();
  for (size_t i = 0; i < ratio_in_order.size(); ++i)
    if (ratio_in_order[i] < 0.5 * r0) return (int)i;
  return (int)ratio_in_order.size();
}

void dump_map(const char* path, const Vector3fVector& pts) {
  FILE* f = std::fopen(path, "w");
  if (!f) return;
  for (const auto& p : pts) std::fprintf(f, "%.9g %.9g %.9g\n", p.x(), p.y(), p.z());
  std::fclose(f);
}

#ifdef VO_B200_DROPIN
// The same loop through the device-resident frame pipeline (vo_pipe_*, include/vo_b200.h §5):
// only the measurements go up and the pose comes down; matches, join, cloud and map stay on the GPU.
int run_pipeline(int n_landmarks, int n_frames, unsigned seed, int rounds, const std::string& dump) {
  Eigen::Matrix3f k;
  k << 180.f, 0.f, 320.f, 0.f, 180.f, 240.f, 0.f, 0.f, 1.f;
  Camera synth_cam(480, 640, 0, 5, k);
  const World world = make_world(n_landmarks, seed);
  Robot robot(seed + 7u);
  Frame reference = observe(world, synth_cam, robot);
  robot.step();
  Frame current = observe(world, synth_cam, robot);

  vo_camera cam;
  cam.rows = 480, cam.cols = 640, cam.z_near = 0, cam.z_far = 5;
  vo_b200::pack3(k, cam.K);
  vo_b200::pack_iso(Eigen::Isometry3f::Identity(), cam.T);
  vo_pipe_t pipe = nullptr;
  vo_b200::check(vo_pipe_create(&pipe, vo_b200::device(), &cam, 32768, (int64_t)n_landmarks + 4096),
                 "vo_pipe_create");
  auto pts = [](Frame& f) { return f.pc.size() ? f.pc.points()[0].data() : nullptr; };
  auto app = [](Frame& f) { return f.pc.size() ? f.pc.appearances()[0].data() : nullptr; };

  auto t0 = Clock::now();
  vo_b200::check(vo_pipe_first_frame(pipe, pts(reference), app(reference), (int64_t)reference.pc.size()),
                 "vo_pipe_first_frame");
  IntPairVector corr(std::min(reference.pc.size(), current.pc.size()) + 1);
  int64_t n_corr = 0;
  vo_b200::check(vo_pipe_second_frame(pipe, pts(current), app(current), (int64_t)current.pc.size(),
                                      &corr[0].first, (int64_t)corr.size(), &n_corr),
                 "vo_pipe_second_frame");
  corr.resize((size_t)n_corr);
  const Eigen::Isometry3f X0 = estimate_transform(k, corr, reference.pc.points(), current.pc.points());
  float Xf[16];
  vo_b200::pack_iso(X0, Xf);
  vo_b200::check(vo_pipe_bootstrap(pipe, Xf), "vo_pipe_bootstrap");
  const Eigen::Isometry3f gt0 = current.world_in_camera * reference.world_in_camera.inverse();
  const float scale = X0.translation().norm() / std::max(1e-12f, gt0.translation().norm());
  Eigen::Isometry3f gt_prev = current.world_in_camera;
  const double t_init = ms_since(t0);

  std::vector<double> rot_err, ratio;
  long long sum_corr = 0, sum_meas = 0;
  FILE* fd = dump.empty() ? nullptr : std::fopen(dump.c_str(), "w");
  double loop_ms = 0;
  int frames_done = 0, overflow = 0;
  for (int f = 2; f < n_frames; ++f) {
    robot.step();
    current = observe(world, synth_cam, robot);  // not timed: stands for the sensor
    const auto tf = Clock::now();
    vo_pipe_result res;
    vo_b200::check(vo_pipe_step(pipe, pts(current), app(current), (int64_t)current.pc.size(), rounds,
                                10000.f, &res),
                   "vo_pipe_step");
    if (f - 2 >= warmup_frames(n_frames)) {  // the first frames pay one-off costs (module load, first touch)
      loop_ms += ms_since(tf);
      ++frames_done;
    }
    overflow |= res.map_overflow;
    sum_corr += res.n_correspondences;
    sum_meas += res.n_measurements;
    Eigen::Isometry3f X_curr = Eigen::Isometry3f::Identity();
    vo_b200::unpack_iso(res.T, X_curr);
    const Eigen::Isometry3f gt = current.world_in_camera * gt_prev.inverse();
    const Eigen::Matrix3f dR = X_curr.linear().transpose() * gt.linear();
    rot_err.push_back(rotation_angle(dR));
    ratio.push_back(X_curr.translation().norm() / std::max(1e-12f, gt.translation().norm()));
    if (fd) {
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) std::fprintf(fd, "%.9g ", X_curr(r, c));
      std::fprintf(fd, "\n");
    }
    gt_prev = current.world_in_camera;
  }
  if (fd) std::fclose(fd);
  // the map (synchronises with the last frame's asynchronous merge; inside the timed total)
  const auto tm = Clock::now();
  Vector3fVector map_pts((size_t)n_landmarks + 4096);
  int64_t n_map = 0;
  vo_b200::check(vo_pipe_get_map(pipe, map_pts[0].data(), nullptr, (int64_t)map_pts.size(), &n_map),
                 "vo_pipe_get_map");
  loop_ms += ms_since(tm);
  map_pts.resize((size_t)n_map);
  if (const char* mp = std::getenv("VO_SEQ_MAPDUMP")) dump_map(mp, map_pts);
  vo_pipe_destroy(pipe);

  double rot_mean = 0;
  for (double e : rot_err) rot_mean += e;
  rot_mean /= std::max<size_t>(1, rot_err.size());
  const int alive = scale_alive_frames(ratio);
  std::sort(ratio.begin(), ratio.end());
  const double ratio_med = ratio.empty() ? 0 : ratio[ratio.size() / 2];
  std::printf(
      "{\"impl\": \"b200-pipeline\", \"landmarks\": %d, \"frames\": %d, \"rounds\": %d, \"seed\": %u, "
      "\"loop_ms\": %.3f, \"frames_per_s\": %.3f, \"init_ms\": %.3f, "
      "\"stage_ms_per_frame\": {\"step\": %.4f}, "
      "\"mean_measurements\": %.1f, \"mean_correspondences\": %.1f, \"map_points\": %lld, "
      "\"map_overflow\": %d, "
      "\"rot_err_mean_rad\": %.3e, \"scale_first_pair\": %.6f, \"scale_median\": %.6f, "
      "\"scale_alive_frames\": %d}\n",
      n_landmarks, frames_done, rounds, seed, loop_ms, frames_done / (loop_ms * 1e-3), t_init,
      loop_ms / frames_done, double(sum_meas) / (n_frames - 2), double(sum_corr) / (n_frames - 2),
      (long long)n_map, overflow, rot_mean, (double)scale, ratio_med, alive);
  return 0;
}
#endif

}  // namespace

int main(int argc, char** argv) {
#ifdef VO_B200_DROPIN
  if (argc == 2 && std::string(argv[1]) == "init") {
    // creates and releases the CUDA context + one solver: what every GPU executable pays once
    PICPSolver warm;
    (void)warm;
    return 0;
  }
#endif
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s n_landmarks n_frames [seed] [rounds] [pose_dump]\n", argv[0]);
    return 2;
  }
  const int n_landmarks = std::atoi(argv[1]), n_frames = std::atoi(argv[2]);
  const unsigned seed = argc > 3 ? (unsigned)std::atoi(argv[3]) : 1000u;
  const int rounds = argc > 4 ? std::atoi(argv[4]) : 100;
  const std::string dump = argc > 5 ? argv[5] : "";
  if (n_frames < 3) {
    std::fprintf(stderr, "need at least 3 frames\n");
    return 2;
  }
#ifdef VO_B200_DROPIN
  // VO_SEQ_MODE=classes: through the drop-in classes (the reference's call surface, one call at a
  // time); default: the device-resident frame pipeline
  const char* mode = std::getenv("VO_SEQ_MODE");
  if (!(mode && std::string(mode) == "classes")) return run_pipeline(n_landmarks, n_frames, seed, rounds, dump);
#endif

  Eigen::Matrix3f k;
  k << 180.f, 0.f, 320.f, 0.f, 180.f, 240.f, 0.f, 0.f, 1.f;
  Camera synth_cam(480, 640, 0, 5, k);  // generates the measurements
  Camera cam(480, 640, 0, 5, k);        // the solver's camera
  const World world = make_world(n_landmarks, seed);
  Robot robot(seed + 7u);

  Frame reference = observe(world, synth_cam, robot);
  robot.step();
  Frame current = observe(world, synth_cam, robot);

  double t_assoc = 0, t_join = 0, t_picp = 0, t_tri = 0, t_map = 0, t_init = 0;
  std::vector<double> rot_err, ratio;
  long long sum_corr = 0, sum_meas = 0;
  FILE* fd = dump.empty() ? nullptr : std::fopen(dump.c_str(), "w");

  // ---- first pair: epipolar initialisation + first triangulation (vo_complete.cpp:116-147) ----
  auto t0 = Clock::now();
  IntPairVector corr_imgs = associate(reference.pc.appearances(), current.pc.appearances());
  const Eigen::Isometry3f X0 =
      estimate_transform(cam.cameraMatrix(), corr_imgs, reference.pc.points(), current.pc.points());
  PointCloudVector<3> triangulated;
  IntPairVector corr_world;
  triangulate_points(k, X0, corr_imgs, reference.pc, current.pc, triangulated, corr_world);
  PointCloudVector<3> map;
  map.update(triangulated);
  Eigen::Isometry3f history = X0.inverse();
  Eigen::Isometry3f X_curr = X0;
  Eigen::Isometry3f gt_prev = current.world_in_camera;
  // monocular scale: fixed by the first pair
  const Eigen::Isometry3f gt0 = current.world_in_camera * reference.world_in_camera.inverse();
  const float scale = X0.translation().norm() / std::max(1e-12f, gt0.translation().norm());
  reference = current;
  t_init = ms_since(t0);

  PICPSolver solver;
  solver.setKernelThreshold(10000);
  double loop_ms = 0;
  int frames_done = 0;
  for (int f = 2; f < n_frames; ++f) {
    robot.step();
    current = observe(world, synth_cam, robot);  // not timed: stands for the sensor
    const bool timed = f - 2 >= warmup_frames(n_frames);  // same warm-up rule as the pipeline mode
    const auto tf = Clock::now();

    auto t = Clock::now();
    corr_imgs = associate(reference.pc.appearances(), current.pc.appearances());
    if (timed) t_assoc += ms_since(t);

    t = Clock::now();
    corr_world = join_with_world(corr_imgs, corr_world);
    const PointCloudVector<3> moved = X_curr * triangulated;
    if (timed) t_join += ms_since(t);

    sum_corr += (long long)corr_world.size();  // the solver's input
    t = Clock::now();
    cam.setWorldInCameraPose(Eigen::Isometry3f::Identity());
    solver.init(cam, moved.points(), current.pc.points());
#ifdef VO_B200_DROPIN
    solver.compute(corr_world, false, rounds);  // all rounds in one resident launch
#else
    for (int i = 0; i < rounds; ++i) solver.oneRound(corr_world, false);
#endif
    cam = solver.camera();
    X_curr = cam.worldInCameraPose();
    if (timed) t_picp += ms_since(t);

    t = Clock::now();
    triangulate_points(k, X_curr, corr_imgs, reference.pc, current.pc, triangulated, corr_world);
    if (timed) t_tri += ms_since(t);

    t = Clock::now();
    map.update(history * triangulated);
    history = history * X_curr.inverse();
    if (timed) t_map += ms_since(t);

    if (timed) {
      loop_ms += ms_since(tf);
      ++frames_done;
    }
    if (std::getenv("VO_SEQ_LOG") && f % 50 == 0)
      std::fprintf(stderr, "frame %d: meas %zu corr %zu map %zu |t| %.4f cum ms: assoc %.1f join %.1f picp %.1f tri %.1f map %.1f\n",
                   f, current.pc.size(), corr_world.size(), map.size(), X_curr.translation().norm(),
                   t_assoc, t_join, t_picp, t_tri, t_map);
    sum_meas += (long long)current.pc.size();

    // accuracy of the relative pose against the generator's ground truth
    const Eigen::Isometry3f gt = current.world_in_camera * gt_prev.inverse();
    const Eigen::Matrix3f dR = X_curr.linear().transpose() * gt.linear();
    rot_err.push_back(rotation_angle(dR));
    ratio.push_back(X_curr.translation().norm() / std::max(1e-12f, gt.translation().norm()));
    if (fd) {
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) std::fprintf(fd, "%.9g ", X_curr(r, c));
      std::fprintf(fd, "\n");
    }
    gt_prev = current.world_in_camera;
    reference = current;
  }
  if (fd) std::fclose(fd);
  if (const char* mp = std::getenv("VO_SEQ_MAPDUMP")) {
    const PointCloudVector<3>& cm = map;
    dump_map(mp, cm.points());
  }

  double rot_mean = 0;
  for (double e : rot_err) rot_mean += e;
  rot_mean /= std::max<size_t>(1, rot_err.size());
  const int alive = scale_alive_frames(ratio);
  std::sort(ratio.begin(), ratio.end());
  const double ratio_med = ratio.empty() ? 0 : ratio[ratio.size() / 2];
#ifdef VO_B200_DROPIN
  const char* impl = "b200";
#else
  const char* impl = "reference-cpu";
#endif
  std::printf(
      "{\"impl\": \"%s\", \"landmarks\": %d, \"frames\": %d, \"rounds\": %d, \"seed\": %u, "
      "\"loop_ms\": %.3f, \"frames_per_s\": %.3f, \"init_ms\": %.3f, "
      "\"stage_ms_per_frame\": {\"associate\": %.4f, \"join_transform\": %.4f, \"picp\": %.4f, "
      "\"triangulate\": %.4f, \"map_update\": %.4f}, "
      "\"mean_measurements\": %.1f, \"mean_correspondences\": %.1f, \"map_points\": %zu, "
      "\"rot_err_mean_rad\": %.3e, \"scale_first_pair\": %.6f, \"scale_median\": %.6f, "
      "\"scale_alive_frames\": %d}\n",
      impl, n_landmarks, frames_done, rounds, seed, loop_ms, frames_done / (loop_ms * 1e-3), t_init,
      t_assoc / frames_done, t_join / frames_done, t_picp / frames_done, t_tri / frames_done,
      t_map / frames_done, double(sum_meas) / (n_frames - 2), double(sum_corr) / (n_frames - 2),
      map.size(), rot_mean, (double)scale, ratio_med, alive);
  return 0;
}
