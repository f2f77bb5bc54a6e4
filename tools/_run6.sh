mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log | grep -v Warning
timeout 600 python tools/sweep.py > gpurun_out/sweep.md 2> gpurun_out/sweep.err; echo "sweep rc=$?"; cat gpurun_out/sweep.md
