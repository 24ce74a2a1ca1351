#!/bin/bash
# N-GPU e2e with the host raster interleaved over the memory nodes (default) and on one node (RZ_HOST_INTERLEAVE=0)
N=${1:-2}
python -m pytest tests/test_gpu_api.py tests/test_abi_c.py tests/test_gpu_multi.py -x -q 2>&1 | tail -3
cat /sys/devices/system/node/online; nproc; for b in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//'); do echo $b $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null); done; grep -i 'cap' /proc/self/status | head -3
for il in 1 0; do
  RZ_HOST_INTERLEAVE=$il python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 3 --warmup 3 --others none > gpurun_out/interleave${il}_n$N.json 2> gpurun_out/interleave${il}_n$N.err
  tail -c 300 gpurun_out/interleave${il}_n$N.err
  python - <<EOF
import json
d=json.loads([l for l in open("gpurun_out/interleave${il}_n$N.json") if l.startswith("{")][-1])
e=d["e2e"]; print("interleave $il N=$N step", d["ms_per_step"], "e2e", e["ms_per_step"], e["ms_each_step"], "flat", e["flatten_ms"], "cached", e["e2e_handle_cached"]["ms_per_step"], e["host_raster_vs_oracle"], (e.get("host_d2h_ceiling") or {}))
for p in e["per_device"]: print(p)
EOF
done
