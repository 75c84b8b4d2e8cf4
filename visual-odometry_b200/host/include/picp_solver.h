// picp_solver.h — drop-in for the reference's include/picp_solver.h (class PICPSolver, :18-79).
//
// Usage is unchanged: init(camera, world_points, image_points), then oneRound(correspondences,
// keep_outliers) as often as wanted, then camera() / chiInliers() / ... .  What changed is WHERE
// it runs: init() uploads the two point sets once; oneRound() only QUEUES a Gauss-Newton iteration,
// and the queue is flushed as ONE kernel launch (projection, 2x6 Jacobians, robust weights, the 6x6
// H / b reduction, the LDL^T solve and the pose update of all queued rounds stay on the device) when
// an accessor is read, the correspondences or keep_outliers change, or init() is called — the
// reference's main calls oneRound() a hundred times and then camera() (vo_complete.cpp:164-166): one
// launch and one synchronisation per frame instead of a hundred.  n rounds in one launch are
// bit-identical to n launches of one round (tests/test_picp_gpu.py).  compute(correspondences,
// keep_outliers, n) is the additive multi-round entry.
#pragma once
#include <vector>

#include "camera.h"
#include "defs.h"
#include "utils.h"

struct vo_picp_s;

class PICPSolver {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW

  PICPSolver();
  ~PICPSolver();
  PICPSolver(const PICPSolver&) = delete;  // owns a device handle
  PICPSolver& operator=(const PICPSolver&) = delete;

  // reference picp_solver.cpp:16-23.  The reference borrows the two vectors; here they are
  // copied to the device, so the caller may change them afterwards without effect until the
  // next init().
  void init(const Camera& camera, const Vector3fVector& world_points,
            const Vector2fVector& image_points);

  inline float kernelThreshold() const { return _kernel_thereshold; }
  void setKernelThreshold(float kernel_threshold);

  // accessors: each one waits for the rounds queued so far (picp_solver.h:41-50)
  const Camera& camera() const;
  const float chiInliers() const;
  const float chiOutliers() const;
  const int numInliers() const;

  // one Gauss-Newton iteration (picp_solver.cpp:98-112); correspondences are
  // (first: measurement index, second: world-point index).  Queued (see above).
  bool oneRound(const IntPairVector& correspondences, bool keep_outliers);
  // `rounds` iterations without leaving the device
  bool compute(const IntPairVector& correspondences, bool keep_outliers, int rounds);

 protected:
  bool upload(const IntPairVector& correspondences);  // true if the device copy had to change
  void flush() const;    // launch the queued rounds
  void refresh() const;  // device state -> host mirrors

  vo_picp_s* _handle;
  mutable Camera _camera;
  float _kernel_thereshold;  // (sic) the reference's spelling, kept for subclass compatibility
  float _damping;
  int _min_num_inliers;
  std::vector<int> _pairs_cache;  // last uploaded correspondences (flattened)
  mutable bool _dirty;            // rounds were launched since the last refresh
  mutable int _pending_rounds;    // oneRound() calls not launched yet
  bool _pending_keep;             // their keep_outliers
  mutable float _chi_inliers, _chi_outliers;
  mutable int _num_inliers;
  // host mirrors of the last linearisation (damping included), refreshed with the accessors above;
  // protected members of the same names exist in the reference (picp_solver.h:70-71) and its
  // subclasses read them
  mutable Matrix6f _H;
  mutable Vector6f _b;
};
