#!/bin/bash
# ncu evidence for one bench step (run on the GPU box via gpurun). Usage: scripts/profile.sh <tag>
# 1) plain run must exit 0; 2) per-launch device times of one step; 3) full-counter captures of the top kernels.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-vq --train-batch= --no-scaled --ragged-stream 0"
$CMD > $OUT/plain_$TAG.log 2> $OUT/plain_$TAG.err || { echo "plain run failed"; tail -n 20 $OUT/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 17 -c 1 -f -o $OUT/attn_$TAG $CMD > $OUT/ncu_attn_$TAG.log 2>&1
echo "attn capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 70 -c 6 -f -o $OUT/gemm_$TAG $CMD > $OUT/ncu_gemm_$TAG.log 2>&1
echo "gemm capture rc=$?"
# quantizer kernel: the K=4096 / D=128 case of the quick microbench (launches 14..26 of vq_argmin2_kernel) and the K=65536 one
python scripts/vq_bench.py --quick > $OUT/plain_vq_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vq_argmin -s 18 -c 1 -f -o $OUT/vq_$TAG python scripts/vq_bench.py --quick > $OUT/ncu_vq_$TAG.log 2>&1
echo "vq capture rc=$?"
ls -la $OUT
