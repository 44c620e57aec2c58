timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_infer_gpu.py -x -q -m gpu -k "graph or pipeline" 2>&1 | tail -6
for g in 0 1; do GRAPH=$g RAW=vitg timeout 600 python tools/bench_pipeline.py 2>&1 | tail -1; done
