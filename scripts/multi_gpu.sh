#!/bin/bash
# Multi-GPU measurements on one box (run under `gpurun --gpus N`): the contract workload with the in-kernel exchanges and with NCCL,
# configs[3] (ScanNet-shaped scene, width 256, sharded by keyframe), ray-batch sweep points (configs[4]) and a strong-scaling
# point at the reference's 5120-ray BA batch.  One JSON line per run in gpurun_out/<tag>.jsonl.
N=${1:-2}
tag=${2:-multi}
out=gpurun_out/${tag}_n${N}.jsonl
: > $out
port=29600
run() {
  port=$((port + 1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --no-extras "$@" 2> gpurun_out/${tag}_n${N}_last.err | grep '^{' >> $out
  tail -2 gpurun_out/${tag}_n${N}_last.err | grep -i -E "error|Traceback" 
}
run
[ -z "$SKIP_NCCL" ] && run --nccl
run --workload scannet_large --width 256 --steps 10
[ $N -lt 8 ] && [ -z "$SKIP_NCCL" ] && run --workload scannet_large --width 256 --steps 10 --nccl
shift 2
for r in "$@"; do run --rays $r --steps 5; done
run --rays 5120 --strong
python - $out <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    c = d["config"]
    print(d["n_gpus"], d["scaling"], c["workload"][:24], "w", c["workload"].split("width ")[1][:3], "rays/gpu", c["rays_per_gpu"], "ms", round(d["ms_per_step"], 4),
          "Mrays/s", round(d["value"] / 1e6, 2), "e2e", round(d["e2e"]["value"] / 1e6, 2), "|", c["parallelism"][:60], "|", (d.get("exchange_check") or {}).get("ok"))
PY
