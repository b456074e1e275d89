#!/bin/bash
# Experiment build: only the <panda, peg-in-hole> kernels of pih_kernels.cu (-DPIH_FAST_BUILD), extra -D flags from the
# command line; the pipe translation unit is taken prebuilt from csrc/build/ (python -c "import __graft_entry__ as g; g.build()").
# usage: tools/fast_build.sh out_name [-DPIH_PHASE_TICKS ...]   ->  tools/_build/libpih_<out_name>.so
set -e
here=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$here/_build"
out="$here/_build/libpih_$1.so"; shift
cd "$here/../peg_in_hole_gym_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -shared -diag-suppress 177 \
  -split-compile 8 -DPIH_FAST_BUILD "$@" -o "$out" pih_kernels.cu build/pih_pipe.o
echo "$out"
