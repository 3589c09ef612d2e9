timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attn" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_reference.py -x -q -m gpu -k "kernels" 2>&1 | tail -3
timeout 120 python scripts/attn_bench.py 64 2>&1 | tail -1
timeout 120 python scripts/attn_bench.py 16 2>&1 | tail -1
timeout 120 python scripts/attn_bench.py 3 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_backward.py -x -q -m gpu 2>&1 | tail -3
