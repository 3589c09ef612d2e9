#!/bin/bash
# full-counter capture of the attention kernel inside one bench step (GPU box). Usage: scripts/profile_attn.sh <tag>
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.log 2> $OUT/plain_$TAG.err || { echo "plain run failed"; tail -n 20 $OUT/plain_$TAG.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 17 -c 1 -f -o $OUT/attn_$TAG $CMD > $OUT/ncu_attn_$TAG.log 2>&1
echo "attn capture rc=$?"
