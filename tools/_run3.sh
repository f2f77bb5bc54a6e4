mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest.log | grep -v Warning
timeout 900 python bench.py --steps 20 --frames 512 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('step ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
for r in d['kernels']['per_stage']: print(r['kernel'], round(r['ms']*1e3,1))
print('sustained', d['sustained']['ms_per_step'])
print('pruned', json.dumps(d.get('pruned'), indent=1))
print('cfg3', json.dumps(d['cfg3'], indent=1))
f=d['frames']; print('frames', {k:f[k] for k in f if k!='workload'})
PY
tail -3 gpurun_out/bench.err
