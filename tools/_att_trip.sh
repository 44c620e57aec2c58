timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k attention 2>&1 | tail -3
for v in 0 1 3 2; do ADA_ATT_VARIANT=$v timeout 60 python tools/bench_attention.py 2>&1 | tail -1; done
ADA_ATT_VARIANT=0 N=5477 B=4 timeout 60 python tools/bench_attention.py 2>&1 | tail -1
