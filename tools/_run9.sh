timeout 600 python -m pytest tests/test_gpu_trace.py -m gpu -q --timeout 120 2>&1 | tail -30 | grep -v Warn
