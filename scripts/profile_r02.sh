#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU, ONE call: all ncu runs of a call count as one). One sample group of 256
# slices per step. 1) plain run (must exit 0 first), 2) launch list (gpu__time_duration), 3) DRAM traffic of every conv
# launch, 4) --set full captures: the halo conv kernel (first forward launches of a step), the flattened-row kernel, the
# weight-gradient kernels.
set -u
export SPFF_BENCH_SAMPLES=256
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-extras"
OUT=gpurun_out
TAG=${1:-r02}
$CMD > $OUT/prof_plain_$TAG.json 2> $OUT/prof_plain_$TAG.err || { echo "plain run failed"; tail -5 $OUT/prof_plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv3_ -c 1200 --csv \
    --log-file $OUT/conv_traffic_$TAG.csv $CMD > $OUT/ncu_traffic.log 2>&1
cap() {  # name, kernel regex, skip, count. The raw page goes back as CSV; the .ncu-rep files together exceed what gpurun
         # copies back (64 MiB), so only the halo capture's report is kept (source page).
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_$1_$TAG $CMD > $OUT/ncu_$1.log 2>&1
  ncu -i $OUT/prof_$1_$TAG.ncu-rep --page raw --csv > $OUT/prof_$1_$TAG.csv 2>/dev/null
  if [ "$1" != "halo" ]; then rm -f $OUT/prof_$1_$TAG.ncu-rep; fi
}
cap halo "^conv3_halo_kernel" 0 6
cap rows "^conv3_fprop_kernel" 0 4
cap wgrad "^conv3_wgrad_kernel" 0 3
cap wgrad_kh "^conv3_wgrad_kh_kernel" 0 2
cap bw_maxpool "^maxpool_bwd_codes" 2 1
cap bw_convt "^pw_gemm_kernel" 0 3
cap bw_head "^head_loss_mma" 0 1
cap stem "^stem_fwd_stats_mma" 0 1
ls -la $OUT/ | head -40; du -sh $OUT
