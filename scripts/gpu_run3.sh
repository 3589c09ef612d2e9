for v in _trace _trace_nomufu; do
  export TTK_LIB_PATH=$PWD/titok_video_b200/lib/libtitok_b200$v.so
  python scripts/attn_bench.py 64 --trace 2>&1 | tail -32
done
