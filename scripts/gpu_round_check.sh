#!/bin/bash
# One GPU-box call: new-path test first, then the whole GPU suite, the bench (default flags), a same-box A/B of the
# latent-tail path, the ncu launch list of one bench step, and the smoke entry. Usage: scripts/gpu_round_check.sh <tag>
TAG=${1:-t1}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "latent_tail or bucketed" -p no:cacheprovider > $OUT/${TAG}_tail.log 2>&1
echo "tail tests rc=$? $(tail -n 1 $OUT/${TAG}_tail.log)"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1
echo "gpu suite rc=$? $(tail -n 1 $OUT/${TAG}_tests.log)"
grep -E "^(FAILED|ERROR)" $OUT/${TAG}_tests.log | head -n 20
timeout 400 python bench.py > $OUT/${TAG}_bench.log 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
QUICK="--no-cpu-baseline --no-gpu-reference --no-vq --train-batch= --no-scaled --ragged-stream 0"
TTK_LATENT_TAIL=0 timeout 200 python bench.py $QUICK > $OUT/${TAG}_bench_notail.log 2> $OUT/${TAG}_bench_notail.err
echo "bench (all rows through the last encoder layer) rc=$?"
timeout 200 python bench.py $QUICK > $OUT/${TAG}_bench_tail.log 2> $OUT/${TAG}_bench_tail.err
echo "bench (latent tail, quick) rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $OUT/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 3 $QUICK > $OUT/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$? $(tail -n 1 $OUT/${TAG}_smoke.log)"
