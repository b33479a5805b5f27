#!/bin/bash
# Round-2 profiling pass (run under gpurun, one GPU): plain bench line, ncu launch list, ncu --set full of the hot kernels at the
# contract workload, the same for the width-256 ScanNet-shaped workload, and the single-GPU ray sweep.
tag=${1:-r02}
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || { tail -5 gpurun_out/${tag}_bench.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_field_bw|k_field_pp|k_intersect_warp|k_sample_warp|k_tri_scatter|k_tri_gather|k_composite_fwd|k_composite_bwd|k_loss_reduce|k_wgrad_finish' --launch-skip 50 -c 10 -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${tag}_ncu_full.log 2>&1
python bench.py --workload scannet_large --width 256 --no-extras --steps 10 > gpurun_out/${tag}_bench_scannet_w256.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:'k_field_w256|k_wgrad_w256' --launch-skip 9 -c 3 -f -o gpurun_out/${tag}_w256_full python bench.py --workload scannet_large --width 256 --steps 2 --warmup 3 --no-extras > gpurun_out/${tag}_ncu_w256.log 2>&1
out=gpurun_out/${tag}_sweep.jsonl
: > $out
for r in 4096 16384 65536 262144 524288; do
  timeout 200 python bench.py --rays $r --steps 5 --warmup 3 --no-extras 2>/dev/null | grep '^{' >> $out
done
ls -la gpurun_out/${tag}_*
