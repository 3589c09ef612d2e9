timeout 900 python -m pytest tests/test_gpu_reference.py -x -q -m gpu -k "packed_discriminator" 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-vq --no-gpu-reference --no-cpu-baseline --no-scaled --ragged-stream 0 --train-batch 3 > gpurun_out/r2_bench4.log 2> gpurun_out/r2_bench4.err; tail -c 600 gpurun_out/r2_bench4.err
