set -x
python -m pytest tests/test_gpu_reference.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_ref_test.log
python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_reference.py 2>&1 | tail -15 > gpurun_out/r2_gpu_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.log 2> gpurun_out/r2_bench1.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2_bench1_ref.log 2>&1
tail -c 3000 gpurun_out/r2_ref_test.log; tail -5 gpurun_out/r2_gpu_tests.log; tail -c 1500 gpurun_out/r2_bench1.err
