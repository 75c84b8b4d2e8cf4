#!/usr/bin/env bash
# Builds libvo_b200 with -D experiment switches (nn_tc.cu, picp.cu, triangulate.cu) into build/variants/<name>.so
# (select one at run time with VO_B200_LIB=<path>).   tools/build_variants.sh "name:-DFLAG ..." ...
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
cd "$ROOT"
make -j4 >/dev/null
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-ffp-contract=off --expt-relaxed-constexpr"
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  nvcc $FLAGS $defs -Xptxas -v -c visual-odometry_b200/csrc/nn_tc.cu -o build/variants/nn_tc_$name.o 2> build/variants/$name.ptxas.log
  nvcc $FLAGS $defs -Xptxas -v -c visual-odometry_b200/csrc/picp.cu -o build/variants/picp_$name.o 2>> build/variants/$name.ptxas.log
  nvcc $FLAGS $defs -fmad=false -Xptxas -v -c visual-odometry_b200/csrc/triangulate.cu -o build/variants/tri_$name.o 2>> build/variants/$name.ptxas.log
  nvcc -shared -o build/variants/$name.so build/lib.o build/stage.o build/nn.o build/comm.o build/variants/nn_tc_$name.o \
       build/variants/picp_$name.o build/variants/tri_$name.o build/pipeline.o -lcudart_static -lpthread -ldl -lrt 2>/dev/null
  grep -E "spill|Used" build/variants/$name.ptxas.log | grep -v " 0 bytes spill stores" | { grep -c spill || true; } | sed "s/^/$name: kernels with spills: /"
done
