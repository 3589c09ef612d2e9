for v in "" _noindex; do
  if [ -z "$v" ]; then unset TTK_LIB_PATH; else export TTK_LIB_PATH=$PWD/titok_video_b200/lib/libtitok_b200$v.so; fi
  echo "variant [$v]"
  timeout 300 python scripts/vq_bench.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l)
        if r['kernel']=='ttk_vq_argmin' and r['D']>=64: print('K',r['K'],'D',r['D'],'ms %.3f'%r['ms'],'burst %.3f'%r['frac_of_tensor_peak'])
"
done
