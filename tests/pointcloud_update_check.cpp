// Host check of the drop-in PointCloudVector::update (hash index) against the reference's
// semantics restated naively (PointCloud.h:52-66: first equal appearance is overwritten, else
// append; appended points take part in later matches).  Exit code 0 and "bad=0" when identical.
#include <cmath>
#include <cstdio>
#include <random>

#include "PointCloud.h"

static void naive_update(Vector3fVector& pts, Vector10fVector& apps, const PointCloudVector<3>& cloud) {
  const Vector3fVector np = cloud.points();
  const Vector10fVector na = cloud.appearances();
  for (size_t i = 0; i < np.size(); ++i) {
    bool found = false;
    for (size_t j = 0; j < apps.size() && !found; ++j)
      if (apps[j] == na[i]) {
        pts[j] = np[i];
        found = true;
      }
    if (!found) {
      pts.push_back(np[i]);
      apps.push_back(na[i]);
    }
  }
}

int main() {
  std::mt19937 rng(5);
  std::uniform_real_distribution<float> u(-1.f, 1.f);
  std::uniform_int_distribution<int> pick(0, 199);
  // a pool of appearances with deliberate traps: duplicates, +0/-0 twins, a NaN
  Vector10fVector pool(200);
  for (auto& a : pool)
    for (int d = 0; d < 10; ++d) a(d) = u(rng);
  pool[10] = pool[3];
  pool[20](4) = 0.f;
  pool[21] = pool[20];
  pool[21](4) = -0.f;
  pool[30](7) = std::nanf("");
  PointCloudVector<3> map;
  Vector3fVector ref_pts;
  Vector10fVector ref_apps;
  int bad = 0;
  for (int round = 0; round < 40; ++round) {
    PointCloudVector<3> cloud;
    const int n = 1 + pick(rng) % 60;
    for (int i = 0; i < n; ++i)
      cloud.push_back(PointCloud<3>(Eigen::Vector3f(u(rng), u(rng), u(rng)), pool[pick(rng)]));
    if (round == 17) map.push_back(PointCloud<3>(Eigen::Vector3f(9, 9, 9), pool[3]));  // bypasses update
    if (round == 17) {
      ref_pts.push_back(Eigen::Vector3f(9, 9, 9));
      ref_apps.push_back(pool[3]);
    }
    if (round == 23) {  // caller edits an appearance in place through the non-const accessor
      map.appearances()[0] = pool[150];
      ref_apps[0] = pool[150];
    }
    map.update(cloud);
    naive_update(ref_pts, ref_apps, cloud);
    const PointCloudVector<3>& cm = map;
    const Vector3fVector gp = cm.points();
    const Vector10fVector ga = cm.appearances();
    if (gp.size() != ref_pts.size()) {
      ++bad;
      continue;
    }
    for (size_t j = 0; j < gp.size(); ++j) {
      if (!(gp[j] == ref_pts[j])) ++bad;
      for (int d = 0; d < 10; ++d) {
        const float x = ga[j](d), y = ref_apps[j](d);
        if (!(x == y) && !(x != x && y != y)) ++bad;
      }
    }
  }
  std::printf("map=%zu bad=%d\n", ref_pts.size(), bad);
  return bad == 0 ? 0 : 1;
}
