// vo_b200_host.h — glue shared by the drop-in headers: error policy and device selection.
#pragma once
#include <cstdio>
#include <cstdlib>

#include "vo_b200.h"

// lets a source file that compiles against both header sets (the reference's and this one) pick
// the additive batched entry points when they exist
#define VO_B200_DROPIN 1

namespace vo_b200 {

// There is no CPU fallback: a failing ABI call is fatal, with the library's message.
inline void check(int rc, const char* what) {
  if (rc == VO_OK) return;
  std::fprintf(stderr, "vo_b200: %s failed (status %d): %s\n", what, rc, vo_last_error());
  std::abort();
}
// CUDA device the host layer computes on (VO_B200_DEVICE, default 0)
inline int device() {
  static const int dev = [] {
    const char* e = std::getenv("VO_B200_DEVICE");
    return e ? std::atoi(e) : 0;
  }();
  return dev;
}
// column-major copies for the ABI
template <class M3>
inline void pack3(const M3& m, float out[9]) {
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) out[j * 3 + i] = m(i, j);
}
template <class Iso>
inline void pack_iso(const Iso& x, float out[16]) {
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) out[j * 4 + i] = (i < 3) ? x(i, j) : (j == 3 ? 1.f : 0.f);
}
template <class Iso>
inline void unpack_iso(const float in[16], Iso& x) {
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 3; ++i) x(i, j) = in[j * 4 + i];
}

}  // namespace vo_b200
