mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x ) > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 300 python tools/_prof_e2e.py > gpurun_out/prof_e2e.log 2>&1; echo "prof rc=$?"; head -50 gpurun_out/prof_e2e.log
