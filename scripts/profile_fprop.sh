#!/bin/bash
# ncu --set full of the first conv3_fprop launches of a step (forward: enc1.2 32->32@L1, enc2.1 32->64@L2, enc2.2 64->64@L2,
# enc3.1 64->128@L3, enc3.2 128->128@L3, bott 128->256@L4) — the tensor-pipe evidence for the final build.
set -u
export SPFF_BENCH_SAMPLES=256
CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
$CMD > gpurun_out/prof_plain2.json 2> gpurun_out/prof_plain2.err || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:conv3_fprop_kernel" -s 0 -c 6 -f -o gpurun_out/prof_fprop_r01f $CMD > gpurun_out/ncu_fprop.log 2>&1
ls -la gpurun_out/prof_fprop_r01f.ncu-rep
