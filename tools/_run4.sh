mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pose.py tests/test_gpu_frames.py -m gpu -q --timeout 600 -x 2>&1 | tail -5
timeout 500 python tools/_prof_init.py 2>&1 | grep -v Warn | head -30
timeout 900 python bench.py --steps 20 --frames 512 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('step ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'])
p=d.get('pruned'); print('pruned ms', p.get('ms_per_step'), 'e2e', p.get('e2e_ms_per_step'))
for r in p['per_stage']: print('  ', r['kernel'], round(r['ms']*1e3,1))
f=d['frames']; print('frames', {k:f[k] for k in f if k!='workload'})
PY
tail -3 gpurun_out/bench.err
