timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "vq or vector" 2>&1 | tail -15
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "scaled_config" --durations=6 2>&1 | tail -25
timeout 600 python -m pytest tests/test_gpu_backward.py -x -q -m gpu 2>&1 | tail -4
