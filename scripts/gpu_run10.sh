timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-scaled > gpurun_out/r2_bench_n2.log 2> gpurun_out/r2_bench_n2.err; tail -c 400 gpurun_out/r2_bench_n2.err
