mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_trace.py -q --timeout 300 -x ) > gpurun_out/pytest_trace.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_trace.log
for gs in 0.08 0.06 0.05; do
echo "== grid stop $gs"
SDFR_TRACE_GRID_STOP=$gs timeout 300 python tools/trace_probe.py 64 256 1024 2>&1 | grep "^trace"
SDFR_TRACE_GRID_STOP=$gs SDFR_TRACE_STATS=1 timeout 300 python tools/trace_probe.py 256 2>&1 | grep "launch" | tail -32 | awk '{printf "%s ", $6} END {print ""}'
done
