timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attn" 2>&1 | tail -2
timeout 120 python scripts/attn_bench.py 64 2>&1 | tail -1
timeout 120 python scripts/attn_bench.py 16 2>&1 | tail -1
TTK_LIB_PATH=$PWD/titok_video_b200/lib/libtitok_b200_trace.so python scripts/attn_bench.py 64 --trace 2>&1 | tail -10
