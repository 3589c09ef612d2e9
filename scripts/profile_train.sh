#!/bin/bash
# ncu evidence for the training step (run on the GPU box via gpurun). Usage: scripts/profile_train.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python scripts/train_step.py --batch 16 --steps 2"
$CMD > $OUT/train_plain_$TAG.log 2> $OUT/train_plain_$TAG.err || { echo "plain run failed"; tail -n 20 $OUT/train_plain_$TAG.err; exit 1; }
# launch list of ~one steady-state step (the first 5 steps are warm-up + timed: skip ~5 * 480 launches incl. torch kernels)
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 480 --csv --log-file $OUT/launches_train_$TAG.csv $CMD > $OUT/ncu_launches_train_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 20 -c 2 -f -o $OUT/attnbwd_$TAG $CMD > $OUT/ncu_attnbwd_$TAG.log 2>&1
echo "attn bwd capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 40 -c 4 -f -o $OUT/wgrad_$TAG $CMD > $OUT/ncu_wgrad_$TAG.log 2>&1
echo "wgrad capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rmsnorm_bwd -s 40 -c 2 -f -o $OUT/rmsbwd_$TAG $CMD > $OUT/ncu_rmsbwd_$TAG.log 2>&1
echo "rmsnorm bwd capture rc=$?"
ls -la $OUT | tail -12
