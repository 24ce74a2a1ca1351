#!/bin/bash
# e2e of the one-shot call with the device's rows cut into 1 / 4 / 8 runs (RZ_ONE_SHOT_RUNS), headline workload only
python -m pytest tests/test_gpu_multi.py tests/test_gpu_api.py tests/test_abi_c.py -x -q 2>&1 | tail -3
for r in ${RUNS:-1 4 8}; do
  RZ_ONE_SHOT_RUNS=$r python bench.py --steps 3 --warmup 3 --others none > gpurun_out/one_shot_runs$r.json 2> gpurun_out/one_shot_runs$r.err
  tail -c 300 gpurun_out/one_shot_runs$r.err
  python - <<EOF
import json
d=json.loads([l for l in open("gpurun_out/one_shot_runs$r.json") if l.startswith("{")][-1])
e=d["e2e"]; print("runs $r", d["ms_per_step"], "e2e", e["ms_per_step"], e["ms_each_step"], "flat", e["flatten_ms"], "cached", e["e2e_handle_cached"]["ms_per_step"], e["host_raster_vs_oracle"], e["checksum"])
for p in e["per_device"]: print(p)
EOF
done
