#!/bin/bash
# Ray-batch sweep (BASELINE.json configs[4]) and the large scene (configs[3]) on one GPU; one JSON line per run.
out=${1:-gpurun_out/sweep.jsonl}
: > $out
for r in 4096 8192 16384 32768 65536 131072 262144 524288; do
  timeout 120 python bench.py --rays $r --steps 5 --warmup 3 --no-extras 2>/dev/null | tail -1 >> $out
done
timeout 120 python bench.py --workload scannet_large --rays 8192 --steps 5 --warmup 3 --no-extras 2>/dev/null | tail -1 >> $out
python - $out <<'PY'
import json, sys
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:100]); continue
    print(d["config"]["workload"][:28], d["config"]["rays_per_gpu"], "samples", d["config"]["samples_per_iter_per_gpu"], "ms", round(d["ms_per_step"], 3), "Mrays/s", round(d["value"] / 1e6, 2), {k: round(v, 3) for k, v in d["stage_ms"].items()})
PY
