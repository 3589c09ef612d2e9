# launch list of the bench command + full captures of the attention and VQ kernels (after plain runs exited 0)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-vq --no-scaled --train-batch "" --ragged-stream 0 > gpurun_out/r2_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-vq --no-scaled --train-batch "" --ragged-stream 0 > gpurun_out/r2_ncu_launches.log 2>&1
python scripts/attn_bench.py 64 > gpurun_out/r2_attn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/attn_r2b python scripts/attn_bench.py 64 > gpurun_out/r2_ncu_attn.log 2>&1
python scripts/vq_bench.py --quick > gpurun_out/r2_vq_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vq_argmin -s 8 -c 2 -f -o gpurun_out/vq_r2 python scripts/vq_bench.py --quick > gpurun_out/r2_ncu_vq.log 2>&1
tail -2 gpurun_out/r2_ncu_launches.log gpurun_out/r2_ncu_attn.log gpurun_out/r2_ncu_vq.log
