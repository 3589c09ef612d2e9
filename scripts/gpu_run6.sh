python scripts/attn_bench.py 64 > gpurun_out/r2_attn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/attn_r2a python scripts/attn_bench.py 64 > gpurun_out/r2_ncu_attn.log 2>&1
tail -3 gpurun_out/r2_attn_plain.log; tail -5 gpurun_out/r2_ncu_attn.log
