#!/bin/bash
# GPU box: poisoned-workspace diagnostics (bucket padding clear on / off), then the GPU suite, a quick bench and the smoke entry
TAG=${1:-t2}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/diag_$TAG.log
for clear in 1 0; do
  TTK_BUCKET_CLEAR_ATT=$clear timeout 150 python scripts/diag/tail_bucket_diag.py bucket_clear$clear big small >> $OUT/diag_$TAG.log 2>&1
  echo "rc=$?" >> $OUT/diag_$TAG.log
done
timeout 150 python scripts/diag/tail_bucket_diag.py eager_tail1 mid big >> $OUT/diag_$TAG.log 2>&1
echo "rc=$?" >> $OUT/diag_$TAG.log
grep -E "^\[|rc=|Error" $OUT/diag_$TAG.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1
echo "gpu suite rc=$? $(tail -n 1 $OUT/${TAG}_tests.log)"
grep -E "^(FAILED|ERROR)" $OUT/${TAG}_tests.log | head -n 20
QUICK="--no-cpu-baseline --no-gpu-reference --no-vq --train-batch= --no-scaled"
timeout 200 python bench.py $QUICK > $OUT/${TAG}_bench_quick.log 2> $OUT/${TAG}_bench_quick.err
echo "bench (quick) rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$? $(tail -n 1 $OUT/${TAG}_smoke.log)"
