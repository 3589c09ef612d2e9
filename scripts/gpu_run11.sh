timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "vq or vector" 2>&1 | tail -3
timeout 600 python scripts/vq_bench.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l)
        if r['kernel']=='ttk_vq_argmin': print('K',r['K'],'D',r['D'],'ms %.3f'%r['ms'],'TF %.0f'%r['tflops_algorithmic'],'burst %.3f'%r['frac_of_tensor_peak'],'mma %.3f'%r['mma_frac_of_tensor_peak'])
        else: print(r['kernel'],r['dtype'],'%.0f GB/s'%r['gbs'],'%.3f'%r['frac_of_hbm_peak'])
    else: print(l.rstrip())
"
