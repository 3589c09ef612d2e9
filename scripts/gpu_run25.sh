for v in "" _notoken; do
  if [ -z "$v" ]; then unset TTK_LIB_PATH; else export TTK_LIB_PATH=$PWD/titok_video_b200/lib/libtitok_b200$v.so; fi
  timeout 120 python scripts/attn_bench.py 64 2>&1 | tail -1
  timeout 120 python scripts/attn_bench.py 16 2>&1 | tail -1
done
