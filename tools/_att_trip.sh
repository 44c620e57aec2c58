timeout 600 python -m pytest tests/test_infer_gpu.py -x -q -m gpu -s 2>&1 | tail -12
