#!/bin/bash
# quick GPU check: the pipeline / field / drop-in / parallel tests, a bench line and a launch list (run under gpurun)
tag=${1:-quick}
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_field.py tests/test_gpu_dropin.py tests/test_gpu_parallel.py -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
python bench.py --no-extras > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > /dev/null 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(l for l in open('gpurun_out/${tag}_launches.csv') if l.startswith('"')))
h=rows[0]; ki,vi=h.index("Kernel Name"),h.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki]].append(float(r[vi].replace(",","")))
    except Exception: pass
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1]))[:18]:
    print(f"{k[:70]:70s} n={len(v):4d} avg={sum(v)/len(v)/1000:8.1f}us")
PY
