for lib in "" sdflabel_b200/libsdfr_tile8.so; do
SDFR_LIB=$lib timeout 600 python bench.py --steps 20 --frames 0 --no-cpu > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "lib=$lib bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v.json').read().strip().splitlines()[-1])
print('step ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'sustained', d['sustained']['ms_per_step'])
print({r['kernel']: round(r['ms']*1e3,1) for r in d['kernels']['per_stage']})
p=d.get('pruned'); print('pruned ms', p.get('ms_per_step'), 'e2e', p.get('e2e_ms_per_step'), 'bit-identical', p.get('params_bit_identical_to_whole_lattice'))
print('cfg3', d['cfg3']['device_ms_per_step'], d['cfg3']['whole_lattice_every_iteration']['device_ms_per_step'])
PY
done
