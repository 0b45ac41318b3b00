#!/bin/bash
# developer helper: extract parity tests, then bench.py's per-kernel table for each value of one environment knob
# usage: scripts/fast_ab.sh TAG VAR v1 v2 ...
tag=$1; var=$2; shift 2
python -m pytest tests/test_gpu_extract.py tests/test_gpu_bench_path.py -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -1 gpurun_out/${tag}_tests.log
for v in "$@"; do
  env $var=$v python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --profile-out gpurun_out/${tag}_$v.json > /dev/null 2> gpurun_out/${tag}.err
  python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_$v.json"))
print("$var=$v", {k: round(x["ms"], 3) for k, x in d["per_kernel"].items()}, "step", round(d["step_ms_profiled"], 3))
PY
done
