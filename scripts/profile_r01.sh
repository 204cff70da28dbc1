#!/bin/bash
# Round-1 ncu evidence (run under gpurun, 1 GPU). One sample group of 256 slices per step.
# 1) launch list (gpu__time_duration), 2) DRAM traffic of every conv launch of the run, 3) --set full captures of the
# level-1 launches of the bandwidth kernels and of the 32-channel weight-gradient kernel.
set -u
export SPFF_BENCH_SAMPLES=256
CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
OUT=gpurun_out
$CMD > $OUT/prof_plain.json 2> $OUT/prof_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_r01f.csv $CMD > $OUT/ncu_launch.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv3_ -c 1200 --csv \
    --log-file $OUT/conv_traffic_r01f.csv $CMD > $OUT/ncu_traffic.log 2>&1
cap() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_$1_r01f $CMD > $OUT/ncu_$1.log 2>&1
}
cap bwd_reduce4 norm_act_bwd_reduce4 0 2
cap bwd_apply4 norm_act_bwd_apply4 0 2
cap norm_act "^norm_act_kernel" 0 3
cap norm_act_pool norm_act_pool_kernel 0 1
cap head_loss head_loss 0 1
cap maxpool_bwd maxpool_bwd_add 2 1
cap wgrad32 "conv3_wgrad_kernel" 0 2
cap stem "stem_" 0 2
ls -la $OUT/*.ncu-rep
