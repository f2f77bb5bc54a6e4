mkdir -p gpurun_out
timeout 300 python tools/trace_diff.py 128 0.6 > gpurun_out/trace_diff.log 2>&1; cat gpurun_out/trace_diff.log | grep -v Warn
