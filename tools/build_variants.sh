#!/usr/bin/env bash
# Builds libvo_b200 with experiment switches of nn_tc.cu into build/variants/<name>.so
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
  nvcc -shared -o build/variants/$name.so build/lib.o build/stage.o build/nn.o build/variants/nn_tc_$name.o build/picp.o \
       build/triangulate.o build/pipeline.o -lcudart_static -lpthread -ldl -lrt 2>/dev/null
  grep -A2 nn_tc_filter build/variants/$name.ptxas.log | tail -1 | sed "s/^/$name: /"
done
