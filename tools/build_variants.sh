#!/bin/bash
# build library variants with different -D flags into tools/variants/<name>.so (experiments only)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants
while [ $# -gt 0 ]; do
  name="$1"; flags="$2"; shift 2
  JDB_NVCC_EXTRA="$flags" python -c "from jdeflate_b200.build import build_cuda; build_cuda(force=True)" > /dev/null
  cp jdeflate_b200/lib/libjdeflate.so tools/variants/$name.so
  echo "built $name: $flags"
done
python -c "from jdeflate_b200.build import build_cuda; build_cuda(force=True)" > /dev/null
