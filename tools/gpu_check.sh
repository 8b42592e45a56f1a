#!/bin/bash
# One GPU-box call: scale parity tests, C2 bench with per-kernel profile, optional C3 bench (arg "c3"), round log of one C2 build.
# Everything lands in gpurun_out/<tag>_*.
tag=${1:-check}; shift
mkdir -p gpurun_out
python -m pytest tests/test_prims.py tests/test_scale_parity.py tests/test_stages.py tests/test_dag.py tests/test_artifacts.py tests/test_mikk.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${tag}_tests.txt
cat gpurun_out/${tag}_tests.txt
CLODB200_BENCH_WORKLOAD=C2 CLODB200_PROFILE_OUT=gpurun_out/${tag}_c2_kernels.csv python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_c2_bench.json 2> gpurun_out/${tag}_c2_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_c2_bench.json"))
print("C2", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
PY
head -16 gpurun_out/${tag}_c2_kernels.csv
for what in "$@"; do
  if [ "$what" = "c3" ]; then
    CLODB200_BENCH_WORKLOAD=C3 CLODB200_PROFILE_OUT=gpurun_out/${tag}_c3_kernels.csv python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_c3_bench.json 2> gpurun_out/${tag}_c3_bench.err
    python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_c3_bench.json"))
print("C3", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"]["value"])
PY
    head -20 gpurun_out/${tag}_c3_kernels.csv
  fi
  if [ "$what" = "rounds" ]; then
    CLODB200_DEBUG_ROUNDS=1 python tools/scale_parity.py grid:1300:5 2> gpurun_out/${tag}_rounds.txt | tail -3
    head -c 3000 gpurun_out/${tag}_rounds.txt
  fi
done
