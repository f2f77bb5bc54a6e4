mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
PROBE_N=1850,3700,7400,58000 SDFR_BAND_NP=32 timeout 300 python tools/band_probe.py pair32 2>&1 | grep "^\[" 
PROBE_N=1850,3700,7400,58000 SDFR_BAND_PAIR=1 timeout 300 python tools/band_probe.py dflt 2>&1 | grep "^\["
python - <<'PY'
import numpy as np
a=np.load('gpurun_out/band_probe_pair32.npz'); b=np.load('gpurun_out/band_probe_dflt.npz')
print('pair32 vs default bit-identical:', all(np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for k in a.files))
PY
timeout 600 python bench.py --steps 20 --frames 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('step ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
for r in d['kernels']['per_stage']: print(r['kernel'], round(r['ms']*1e3,1))
print('sustained', d['sustained']['ms_per_step'], 'cfg3', d['cfg3'])
PY
tail -3 gpurun_out/bench.err
