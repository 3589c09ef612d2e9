#!/bin/bash
# ncu evidence for the bandwidth-bound kernels of the inference path (achieved HBM GB/s). Usage: scripts/profile_rows.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-batch= --no-scaled --ragged-stream 0"
$CMD > $OUT/rows_plain_$TAG.log 2> $OUT/rows_plain_$TAG.err || { echo "plain run failed"; tail -n 20 $OUT/rows_plain_$TAG.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"patchify_kernel|embed_kernel|hist_smem|enc_head_fsq|clip_error|normalize_u8" -s 12 -c 10 -f -o $OUT/rows_$TAG $CMD > $OUT/ncu_rows_$TAG.log 2>&1
echo "row kernels capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"fsq_fwd_kernel" -c 2 -f -o $OUT/fsq_$TAG $CMD > $OUT/ncu_fsq_$TAG.log 2>&1
echo "fsq capture rc=$?"
