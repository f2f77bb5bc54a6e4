mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log | grep -v Warning
timeout 300 python tools/_prof_small.py 2>&1 | grep -v Warn
