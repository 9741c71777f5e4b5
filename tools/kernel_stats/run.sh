#!/bin/bash
# Offline (no GPU) statistics of ONE instantiation of felsenstein_walk: registers, spills, SASS size,
# and opcode histograms of the op loops.  Two seconds per variant instead of a 90 s library build.
#   tools/kernel_stats/run.sh <tag> [-DI_K=4 -DI_C=2 -DI_NE=3 -DI_DYN=false -DI_ACCG=false] [other -D switches]
# Outputs build_exp/<tag>.cubin, build_exp/<tag>.sass and prints ptxas' summary + the loop histograms
# (python tools/kernel_stats/loopstat.py build_exp/<tag>.sass).  After a GPU capture,
#   ncu -i X.ncu-rep --page source --csv > src.csv; python tools/kernel_stats/ncu_srcstat.py src.csv 40
# lists the instructions with the most stall samples.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
TAG=$1; shift
mkdir -p "$ROOT/build_exp"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I"$ROOT/mcphylo.jl_b200/csrc" \
     -DKW="\"$ROOT/mcphylo.jl_b200/csrc/kernel_walk.cuh\"" -Xptxas -v -cubin -o "$ROOT/build_exp/$TAG.cubin" "$@" \
     "$HERE/one_kernel.cu" 2>&1 | grep -E "spill|Used|error" || true
cuobjdump -sass "$ROOT/build_exp/$TAG.cubin" > "$ROOT/build_exp/$TAG.sass"
python "$HERE/loopstat.py" "$ROOT/build_exp/$TAG.sass"
