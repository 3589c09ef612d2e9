timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-scaled > gpurun_out/r2_bench_n2b.log 2> gpurun_out/r2_bench_n2b.err; tail -c 600 gpurun_out/r2_bench_n2b.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n2b.log').read().strip().splitlines()[-1])
c=d['config']
print('N=2 value %.0f e2e %.0f tok %.0f' % (d['value'], d['e2e']['value'], d['e2e']['tokens_only']['value']), d['e2e']['copy_probe'])
for k in ('train_step_batch3','train_step_batch3_graph','train_step_batch16','train_step_batch16_graph'):
    v=c.get(k); print(k, v and {kk: v[kk] for kk in ('ms_per_step','clips_per_s','graph_replays','loss')})
print([ (r['K'],r['D'],round(r['frac_burst'],3)) for r in d['roofline']['vq']])
PY
