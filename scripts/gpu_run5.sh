./scripts/micro/pipe_rate
timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "batch_composition" 2>&1 | tail -40
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
