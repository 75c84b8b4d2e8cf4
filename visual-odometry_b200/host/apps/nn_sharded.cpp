// nn_sharded.cpp — a C++ caller of the multi-GPU C ABI (include/vo_b200.h, section 1b): the
// query-sharded appearance nearest neighbour of BASELINE config 4 without Python or torch.
//   nn_sharded <n_gpus> [map_rows=1000000] [queries=100000]
// Builds a uniform 10-D map on the host, plants the queries (3 of 4 are copies of map rows, 1 of 4
// fresh), answers them on 1 GPU and on n GPUs through vo_comm_*, checks that both agree and that
// every planted query finds its row, and prints one JSON line with the timings.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vo_b200.h"

namespace {
uint32_t hash32(uint32_t x) {
  x = ((x >> 16) ^ x) * 0x45d9f3bu;
  x = ((x >> 16) ^ x) * 0x45d9f3bu;
  return (x >> 16) ^ x;
}
float unit(uint32_t h) { return (float)(h >> 8) * (1.0f / 8388608.0f) - 1.0f; }  // [-1, 1)
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define CHECK(expr)                                                         \
  do {                                                                      \
    int rc_ = (expr);                                                       \
    if (rc_ != 0) {                                                         \
      std::fprintf(stderr, "%s -> %d: %s\n", #expr, rc_, vo_last_error()); \
      return 1;                                                             \
    }                                                                       \
  } while (0)
}  // namespace

int main(int argc, char** argv) {
  const int n_gpus = argc > 1 ? std::atoi(argv[1]) : 1;
  const int64_t M = argc > 2 ? std::atoll(argv[2]) : 1000000;
  const int64_t Q = argc > 3 ? std::atoll(argv[3]) : 100000;
  std::vector<float> map((size_t)M * 11), q((size_t)Q * 11);
  for (int64_t r = 0; r < M; ++r) {
    map[(size_t)r * 11] = (float)r;  // the id column, carried and ignored (vo_complete.cpp:22)
    for (int k = 0; k < 10; ++k) map[(size_t)r * 11 + 1 + k] = unit(hash32((uint32_t)(r * 10 + k) + 0x9e3779b9u));
  }
  std::vector<int64_t> target((size_t)Q, -1);
  for (int64_t i = 0; i < Q; ++i) {
    q[(size_t)i * 11] = (float)i;
    if (i % 4 != 3) {
      const int64_t row = (int64_t)(hash32((uint32_t)i * 7919u + 12345u) % (uint32_t)M);
      target[(size_t)i] = row;
      for (int k = 0; k < 10; ++k) q[(size_t)i * 11 + 1 + k] = map[(size_t)row * 11 + 1 + k];
    } else {
      for (int k = 0; k < 10; ++k) q[(size_t)i * 11 + 1 + k] = unit(hash32((uint32_t)(i * 10 + k) + 0x51ed270bu));
    }
  }
  std::vector<int32_t> one((size_t)Q), many((size_t)Q);
  double t_map[2] = {0, 0}, t_query[2] = {0, 0};
  for (int pass = 0; pass < 2; ++pass) {
    const int g = pass == 0 ? 1 : n_gpus;
    vo_comm_t c = nullptr;
    CHECK(vo_comm_init_all(&c, g));
    double t0 = now_ms();
    CHECK(vo_nn_set_map_replicated(c, map.data(), M, 11, 1));
    t_map[pass] = now_ms() - t0;
    std::vector<int32_t>& out = pass == 0 ? one : many;
    CHECK(vo_nn_best_match_sharded(c, q.data(), Q, 11, 0.1f, out.data()));  // warm-up (module load)
    t0 = now_ms();
    const int reps = 3;
    for (int r = 0; r < reps; ++r) CHECK(vo_nn_best_match_sharded(c, q.data(), Q, 11, 0.1f, out.data()));
    t_query[pass] = (now_ms() - t0) / reps;
    CHECK(vo_comm_destroy(c));
  }
  int64_t differ = 0, planted_bad = 0;
  for (int64_t i = 0; i < Q; ++i) {
    differ += one[(size_t)i] != many[(size_t)i];
    // duplicates of a row are possible: the lowest index wins, so the answer is <= the planted row
    if (target[(size_t)i] >= 0 ? (many[(size_t)i] < 0 || many[(size_t)i] > target[(size_t)i]) : many[(size_t)i] != -1)
      ++planted_bad;
  }
  std::printf(
      "{\"app\": \"nn_sharded\", \"n_gpus\": %d, \"map_rows\": %lld, \"queries\": %lld, "
      "\"set_map_ms\": {\"1\": %.3f, \"n\": %.3f}, \"best_match_ms\": {\"1\": %.3f, \"n\": %.3f}, "
      "\"queries_per_s\": {\"1\": %.1f, \"n\": %.1f}, \"sharded_differs_from_single\": %lld, "
      "\"planted_wrong\": %lld}\n",
      n_gpus, (long long)M, (long long)Q, t_map[0], t_map[1], t_query[0], t_query[1], Q / (t_query[0] * 1e-3),
      Q / (t_query[1] * 1e-3), (long long)differ, (long long)planted_bad);
  return (differ == 0 && planted_bad == 0) ? 0 : 3;
}
