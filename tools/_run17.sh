mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trace.py -m gpu -q --timeout 300 -x > gpurun_out/pytest_trace.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_trace.log
timeout 300 python tools/trace_probe.py 64 128 256 512 1024 > gpurun_out/trace_probe.log 2>&1; cat gpurun_out/trace_probe.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/trace_launches.csv python tools/trace_probe.py 256 > gpurun_out/ncu_trace.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/trace_launches.csv')) if len(r) > 10 and r[0].isdigit()]
last = rows[-45:]
tot = 0
for r in last:
    print(r[4][:48], r[-1]); 
PY
