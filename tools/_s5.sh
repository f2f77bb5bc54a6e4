mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_trace.py -q --timeout 300 -x ) > gpurun_out/pytest_trace.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_trace.log
timeout 600 python tools/_trace_bench.py > gpurun_out/trace_bench.json 2> gpurun_out/trace_bench.err; echo "rc=$?"; tail -3 gpurun_out/trace_bench.err; cat gpurun_out/trace_bench.json
