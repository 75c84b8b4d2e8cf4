#!/usr/bin/env bash
# Builds the reference's mains UNCHANGED (compiled where they lie under /root/reference) against
# this repo's drop-in headers + libvo_b200.so:  vo_complete, picp_test, whole_test.
# Every header the mains include resolves to host/include (ours); the reference's include/ is
# NOT on the search path.  The two translation units outside the hot path (file I/O, epipolar
# initialisation) are compiled from the reference's src/ against our headers.
# Outputs go to host/bin (git-ignored; they travel to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${VO_REFERENCE_DIR:-/root/reference}"
[ -d "$REF/src" ] || { echo "build_dropin.sh: $REF not present, keeping prebuilt $HERE/bin" >&2; exit 0; }
mkdir -p "$HERE/bin" "$HERE/obj"
CXX="${CXX:-g++}"
FLAGS="-std=c++17 -O3 -DNDEBUG -w -I $HERE/include -I $ROOT/include -I $ROOT/third_party/mini_eigen"
for f in "$HERE/src/camera.cpp" "$HERE/src/picp_solver.cpp" "$HERE/src/utils.cpp" \
         "$REF/src/files_utils.cpp" "$REF/src/epipolar_utils.cpp"; do
  $CXX $FLAGS -c "$f" -o "$HERE/obj/$(basename "${f%.cpp}").o"
done
OBJS="$HERE/obj/camera.o $HERE/obj/picp_solver.o $HERE/obj/utils.o $HERE/obj/files_utils.o $HERE/obj/epipolar_utils.o"
LINK="-L $ROOT/visual-odometry_b200/lib -lvo_b200 -Wl,-rpath,\$ORIGIN/../../lib"
$CXX $FLAGS "$REF/src/apps/vo_complete.cpp" $OBJS $LINK -o "$HERE/bin/vo_complete"
$CXX $FLAGS "$REF/src/tests/picp_solver_test.cpp" $OBJS $LINK -o "$HERE/bin/picp_test"
$CXX $FLAGS "$REF/src/tests/essential_picp_test.cpp" $OBJS $LINK -o "$HERE/bin/whole_test"
# the synthetic-sequence driver (config 5): our own source, the same file the CPU reference build uses
$CXX $FLAGS "$HERE/apps/vo_sequence.cpp" $OBJS $LINK -o "$HERE/bin/vo_sequence"
# the seeded whole_test driver (config 2), same arrangement
$CXX $FLAGS "$HERE/apps/whole_synthetic.cpp" $OBJS $LINK -o "$HERE/bin/whole_synthetic"
echo "built $HERE/bin/{vo_complete,picp_test,whole_test,vo_sequence} against libvo_b200.so"
