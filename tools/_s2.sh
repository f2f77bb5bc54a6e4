mkdir -p gpurun_out
timeout 300 python tools/trace_probe.py 64 256 1024 > gpurun_out/trace_probe.log 2>&1; tail -4 gpurun_out/trace_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/trace_launches.csv python tools/trace_probe.py 256 > gpurun_out/trace_ncu.log 2>&1; echo "ncu rc=$?"
