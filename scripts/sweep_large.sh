#!/bin/bash
# 1-GPU end of the ray sweep (BASELINE.json configs[4]): 2^20 rays in one launch (132 GiB workspace at 32 samples of capacity per ray),
# 2^21 and 2^22 as 2 / 4 chunks of 2^20 through parallel.ChunkedStep.  One JSON line per run.
out=${1:-gpurun_out/sweep_large.jsonl}
: > $out
timeout 200 python bench.py --rays 8192 --chunks 2 --steps 5 --warmup 3 --no-extras 2>gpurun_out/sweep_large.err | tail -1 >> $out
timeout 300 python bench.py --rays 1048576 --samples-per-ray 32 --steps 3 --warmup 3 --no-extras 2>>gpurun_out/sweep_large.err | tail -1 >> $out
timeout 300 python bench.py --rays 2097152 --samples-per-ray 32 --chunks 2 --steps 3 --warmup 3 --no-extras 2>>gpurun_out/sweep_large.err | tail -1 >> $out
timeout 400 python bench.py --rays 4194304 --samples-per-ray 32 --chunks 4 --steps 3 --warmup 3 --no-extras 2>>gpurun_out/sweep_large.err | tail -1 >> $out
tail -5 gpurun_out/sweep_large.err
python - $out <<'PY'
import json, sys
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:200]); continue
    c = d["config"]
    print(c["rays_per_gpu"], "chunks", c.get("chunks", 1), "samples", c["samples_per_iter_per_gpu"], "ms", round(d["ms_per_step"], 3), "Mrays/s", round(d["value"] / 1e6, 2),
          "e2e", round(d["e2e"]["value"] / 1e6, 2), "loss", c["loss"], "frac", round(d["roofline"]["frac"], 4), "model", round(d["iteration_model"]["ratio"], 3))
PY
