#!/usr/bin/env bash
# Build libpmvs_b200.so with extra nvcc flags (e.g. a -D switch while a kernel variant is being developed), run the GPU
# parity suite and a short bench, then restore the default build.  For the GPU box:
#   gpurun -- 'bash tests/try_experimental.sh -DSOME_SWITCH=1'
# or, to keep nvcc time off the GPU box, build a side copy in the build container first and pass it:
#   PMVS_OUT=$PWD/altlibs/variant.so bash pointmvsnet_b200/csrc/build.sh -DSOME_SWITCH=1
#   gpurun -- 'bash tests/try_experimental.sh altlibs/variant.so'
# Variants that exist at run time are compared without rebuilding: PMVS_OPTIONS="edge=0,gemm=1" (include/pmvs_b200.h)
# and  python tests/profile_per_launch.py C2 "edge=1" "edge=0"  (per-launch CUDA-event times of one pass).
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
