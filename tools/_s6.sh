mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_trace.py -q --timeout 300 -x ) > gpurun_out/pytest_trace.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_trace.log
SDFR_TRACE_STATS=2 timeout 300 python tools/trace_probe.py 1024 2>&1 | grep "launch\|newton" | tail -36 | awk '{printf "%s/%s ", $6, $8} END {print ""}'
timeout 600 python tools/_trace_bench.py 2>/dev/null | grep -A8 "views_in_flight\|\"resolution\|fwd_ms"  | grep "resolution\|fwd_ms\|\"ms\|fwd_rays_per_s\|frac" 
