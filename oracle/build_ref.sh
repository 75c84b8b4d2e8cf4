#!/usr/bin/env bash
# Builds oracle/_ref: the REFERENCE'S OWN sources, compiled where they lie under /root/reference
# (never copied), with the reference's flags (-std=c++17 -O3 -DNDEBUG, no -march: CMakeLists.txt:7)
# against third_party/mini_eigen (Eigen3 is not installed here).  Outputs only into
# oracle/_ref/ (git-ignored, NOT gpurun-ignored: the prebuilt files travel to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${VO_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF/src" ] || { echo "build_ref.sh: $REF not present, keeping prebuilt $OUT" >&2; exit 0; }
mkdir -p "$OUT/bin"
CXX="${CXX:-g++}"
FLAGS="-std=c++17 -O3 -DNDEBUG -ffp-contract=off -fPIC -w -I $HERE/../third_party/mini_eigen -I $REF/include"
LIBSRC="$REF/src/picp_solver.cpp $REF/src/camera.cpp $REF/src/utils.cpp $REF/src/epipolar_utils.cpp $REF/src/files_utils.cpp"
for f in $LIBSRC; do
  o="$OUT/$(basename "${f%.cpp}").o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ "$HERE/../third_party/mini_eigen/Eigen/Core" -nt "$o" ]; then
    $CXX $FLAGS -c "$f" -o "$o"
  fi
done
OBJS="$OUT/picp_solver.o $OUT/camera.o $OUT/utils.o $OUT/epipolar_utils.o $OUT/files_utils.o"
$CXX $FLAGS -shared "$HERE/ref_harness.cpp" $OBJS -o "$OUT/libvo_ref.so"
# the reference's executables, unmodified, as the CPU side of config 1/2/3 (src/CMakeLists.txt:4-8)
$CXX $FLAGS "$REF/src/apps/vo_complete.cpp" $OBJS -o "$OUT/bin/vo_complete"
$CXX $FLAGS "$REF/src/tests/picp_solver_test.cpp" $OBJS -o "$OUT/bin/picp_test"
$CXX $FLAGS "$REF/src/tests/essential_picp_test.cpp" $OBJS -o "$OUT/bin/whole_test"
$CXX $FLAGS "$REF/src/apps/evaluate.cpp" "$REF/src/evaluation_utils.cpp" $OBJS -o "$OUT/bin/evaluation"
# config 5 on the CPU: this repository's sequence driver compiled against the REFERENCE's headers and
# objects (it only uses declarations both header sets share), i.e. the reference implementation of
# every stage on the synthetic frames
$CXX $FLAGS "$HERE/../visual-odometry_b200/host/apps/vo_sequence.cpp" $OBJS -o "$OUT/bin/vo_sequence"
# config 2 on the CPU: the seeded whole_test driver, same arrangement
$CXX $FLAGS "$HERE/../visual-odometry_b200/host/apps/whole_synthetic.cpp" $OBJS -o "$OUT/bin/whole_synthetic"
echo "built $OUT/libvo_ref.so and $OUT/bin/{vo_complete,picp_test,whole_test,evaluation}"
