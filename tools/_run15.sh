mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log | grep -v Warning
timeout 900 python bench.py --steps 20 --frames 512 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('step ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'sustained', d['sustained']['ms_per_step'], 'launches', d['gpu_launches'])
print({r['kernel']: round(r['ms']*1e3,1) for r in d['kernels']['per_stage']})
p=d.get('pruned'); print('pruned ms', p.get('ms_per_step'), 'e2e', p.get('e2e_ms_per_step'), 'bit-identical', p.get('params_bit_identical_to_whole_lattice'))
c=d['cfg3']; print('cfg3', c['device_ms_per_step'], c['whole_lattice_every_iteration']['device_ms_per_step'], c['results_bit_identical_to_whole_lattice'])
f=d['frames']; print('frames', {k:f[k] for k in f if k!='workload'})
print('roofline', d['roofline']['frac'], d['sustained']['lattice_kernel']['frac'])
PY
tail -2 gpurun_out/bench.err
timeout 600 python tools/sweep.py > gpurun_out/sweep.md 2> gpurun_out/sweep.err; echo "sweep rc=$?"; cat gpurun_out/sweep.md
