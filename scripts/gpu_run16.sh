timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for pdl in 0 1 0 1; do
  TTK_PDL=$pdl python bench.py --steps 20 --warmup 5 --no-vq --no-gpu-reference --no-cpu-baseline --no-scaled --train-batch 3,16 > gpurun_out/r2_bench_pdl$pdl.log 2> gpurun_out/r2_bench_pdl$pdl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_pdl$pdl.log').read().strip().splitlines()[-1])
c=d['config']
print('PDL=$pdl value %.0f ms %.3f e2e %.0f tok %.0f | train3 %.3f ms train16 %.3f ms gan %.3f | ragged wall %.3f kernel %.3f ratio %.2f | train ragged %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['tokens_only']['value'], c['train_step_batch3']['ms_per_step'], c['train_step_batch16']['ms_per_step'], c['train_step_gan']['packed']['ms_per_step'], c['ragged_stream']['ms_per_step_wall'], c['ragged_stream']['kernel_ms_per_step'], c['ragged_stream']['wall_over_kernel'], c['train_step_ragged']['ms_per_step_wall']))
PY
done
