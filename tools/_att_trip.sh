RAW=vitg timeout 600 python tools/bench_pipeline.py 2>&1 | tail -2
