start=$(date +%s)
python bench.py > gpurun_out/r2_bench_final.log 2> gpurun_out/r2_bench_final.err; echo "rc=$? bench took $(( $(date +%s) - start )) s"; tail -c 300 gpurun_out/r2_bench_final.err
start=$(date +%s)
python bench.py --impl reference > gpurun_out/r2_bench_final_ref.log 2>&1; echo "rc=$? reference arm took $(( $(date +%s) - start )) s"
