for v in "" _qkvws; do
  if [ -z "$v" ]; then unset TTK_LIB_PATH; else export TTK_LIB_PATH=$PWD/titok_video_b200/lib/libtitok_b200$v.so; fi
  timeout 120 python scripts/gemm_bench.py 64 2>&1 | tail -4
done
export TTK_LIB_PATH=$PWD/titok_video_b200/lib/libtitok_b200_qkvws.so
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm_qkv" 2>&1 | tail -3
