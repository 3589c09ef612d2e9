timeout 600 python -m pytest tests/test_gpu_backward.py -x -q -m gpu -k "graphed" 2>&1 | tail -25
python bench.py --steps 20 --warmup 5 --no-vq --no-gpu-reference --no-cpu-baseline --no-scaled --ragged-stream 0 > gpurun_out/r2_bench5.log 2> gpurun_out/r2_bench5.err; tail -c 800 gpurun_out/r2_bench5.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench5.log').read().strip().splitlines()[-1])
c=d['config']
for k in ('train_step_batch3','train_step_batch3_graph','train_step_batch16','train_step_batch16_graph'):
    v=c.get(k); print(k, v and {kk: v[kk] for kk in ('ms_per_step','wall_ms_per_step','clips_per_s','frac_of_tensor_peak','graph_replays','loss')})
print(d['roofline'].get('stress_init'))
PY
