timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -15
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-vq --no-gpu-reference --no-cpu-baseline > gpurun_out/r2_bench3.log 2> gpurun_out/r2_bench3.err; tail -c 300 gpurun_out/r2_bench3.err
