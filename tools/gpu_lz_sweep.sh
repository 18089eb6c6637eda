#!/bin/bash
# GPU experiment: lz_kernel variants on the mixed and text corpora.  Every argument is one
# configuration: a quoted string of environment assignments (JDB_LZ_*, JDB200_LIB=tools/variants/x.so,
# LEVELS=6, KINDS=5,0, MIB=256)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/lz_sweep.log
: > $out
for cfg in "$@"; do
  echo "== $cfg" >> $out
  ( export $cfg; python tools/gpu_deflate.py ${MIB:-256} ${LEVELS:-6} ${KINDS:-5,0} >> $out 2>&1 )
done
