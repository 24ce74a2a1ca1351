for t in 4 8 16 32; do echo "c4 apply_tiles=$t"; RZ_APPLY_TILES=$t python bench.py --workload c4 --others none --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms_max_over_ranks'].items()}, d['parity_vs_oracle']['bit_exact'])"; done
for t in 1 2 4; do echo "c1 apply_tiles=$t"; RZ_APPLY_TILES=$t python bench.py --workload c1 --others none --steps 20 --warmup 5 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms_max_over_ranks'].items()}, d['parity_vs_oracle']['bit_exact'])"; done
for t in 2 4 8; do echo "c3 apply_tiles=$t"; RZ_APPLY_TILES=$t python bench.py --workload c3 --others none --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms_max_over_ranks'].items()}, d['parity_vs_oracle']['bit_exact'])"; done
