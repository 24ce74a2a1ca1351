#!/bin/bash
# final bench of a round at N GPUs (N = $1): the same command the driver runs, output kept under gpurun_out/
N=${1:-1}
if [ "$N" = 1 ]; then
  python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/final_n$N.json 2> gpurun_out/final_n$N.err
fi
tail -c 400 gpurun_out/final_n$N.err
python - <<EOF
import json
d=json.loads([l for l in open("gpurun_out/final_n$N.json") if l.startswith("{")][-1])
e=d["e2e"]; print("N=$N step", d["ms_per_step"], "frac", d["roofline"]["frac"], "parity", d.get("parity_vs_oracle",{}).get("bit_exact"), "e2e", e["ms_per_step"], e["ms_each_step"], "flat", e["flatten_ms"], "cached", e["e2e_handle_cached"]["ms_per_step"], e["host_raster_vs_oracle"], (e.get("host_d2h_ceiling") or {}).get("ms"))
for p in e["per_device"]: print(p)
for k,v in d["other_configs"].items(): print(k, v.get("ms_per_step"), v.get("parity_vs_oracle",{}).get("bit_exact"), (v.get("e2e") or {}).get("ms_per_step"), (v.get("e2e") or {}).get("host_raster_vs_oracle"))
EOF
