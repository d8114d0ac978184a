#!/usr/bin/env bash
# Build libpmvs_b200.so with experimental compile flags, run the GPU parity suite and a short bench,
# then restore the default build.  For the GPU box:
#   gpurun -- 'bash tests/try_experimental.sh -DPMVS_EDGE_TILE=1'
#   gpurun -- 'bash tests/try_experimental.sh -DPMVS_F32X2=1'
# or, to keep nvcc time off the GPU box, build a side copy in the build container first and pass it:
#   PMVS_OUT=$PWD/altlibs/tile.so bash pointmvsnet_b200/csrc/build.sh -DPMVS_EDGE_TILE=1
#   gpurun -- 'bash tests/try_experimental.sh altlibs/tile.so'
# Flags are described in pointmvsnet_b200/csrc/common.cuh and DESIGN.md section 8.
set -uo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
mkdir -p gpurun_out
tag="$(echo "$*" | tr -c 'A-Za-z0-9' '_')"
if [[ "${1:-}" == *.so ]]; then
  cp pointmvsnet_b200/libpmvs_b200.so gpurun_out/libpmvs_default.so
  cp "$1" pointmvsnet_b200/libpmvs_b200.so
else
  bash pointmvsnet_b200/csrc/build.sh "$@" > "gpurun_out/build_${tag}.log" 2>&1 || { tail -20 "gpurun_out/build_${tag}.log"; exit 1; }
fi
timeout 900 python -m pytest ${PMVS_TESTS:-tests} -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > "gpurun_out/bench_${tag}.json"
python - "$tag" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1]))
print("iters/s", d["value"], "e2e", d["e2e"]["value"])
for k, v in d["kernels"].items():
    print("  %-18s %.4f ms/pass" % (k, v["ms_per_pass"]))
PY
if [[ "${1:-}" == *.so ]]; then
  mv gpurun_out/libpmvs_default.so pointmvsnet_b200/libpmvs_b200.so && echo "default library restored"
else
  bash pointmvsnet_b200/csrc/build.sh > /dev/null 2>&1 && echo "default build restored"
fi
