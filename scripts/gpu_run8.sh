timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 120 python scripts/attn_bench.py 64 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench2.log 2> gpurun_out/r2_bench2.err; tail -c 600 gpurun_out/r2_bench2.err
