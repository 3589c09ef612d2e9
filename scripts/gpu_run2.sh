timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attn" 2>&1 | tail -3
for v in "" _noemu _emu12 _fixed _nosleep; do
  if [ -z "$v" ]; then unset TTK_LIB_PATH; else export TTK_LIB_PATH=$PWD/titok_video_b200/lib/libtitok_b200$v.so; fi
  timeout 120 python scripts/attn_bench.py 64 2>&1 | tail -1
  timeout 120 python scripts/attn_bench.py 16 2>&1 | tail -1
done
export TTK_LIB_PATH=$PWD/titok_video_b200/lib/libtitok_b200_trace.so
timeout 120 python scripts/attn_bench.py 64 --trace 2>&1 | tail -11
unset TTK_LIB_PATH
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
