start=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.log 2> gpurun_out/r2_bench_n8.err; echo "rc=$? took $(( $(date +%s) - start )) s"
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n8.err | tail -5
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n8.log').read().strip().splitlines()[-1])
c=d['config']
print('N=8 value %.0f (%.3f ms) e2e %.0f tok %.0f u8 %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['tokens_only']['value'], d['e2e']['u8']['value']), d['e2e']['copy_probe'])
for k in ('train_step_batch3','train_step_batch3_graph','train_step_batch16','train_step_batch16_graph'):
    v=c.get(k); print(k, v and {kk: v[kk] for kk in ('ms_per_step','ms_per_step_mean','clips_per_s','graph_replays','loss')})
print(c['scaled_config']['clips_per_s'], c['host_affinity'], d['clocks'])
PY
