mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trace.py -m gpu -q --timeout 300 > gpurun_out/pytest_trace.log 2>&1; echo "pytest rc=$?"; grep -n "^E  *Assert\|^E  *assert\|passed\|failed" gpurun_out/pytest_trace.log | head -30
timeout 300 python tools/trace_probe.py 64 128 256 512 1024 2>&1 | grep trace
timeout 900 python bench.py --steps 20 --frames 512 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('step ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'])
f=d['frames']; print('frames', {k:f[k] for k in f if k!='workload'})
print('trace', d.get('trace'))
PY
tail -2 gpurun_out/bench.err
