timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 5 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attn_varlen or (vq_argmin and not microbench) or vector_quantizer or patchify" > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?"
tail -15 gpurun_out/r2_memcheck.log
