mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trace.py -m gpu -q --timeout 300 > gpurun_out/pytest_trace.log 2>&1; echo "pytest rc=$?"; grep -n "^E  *Assert\|^E  *assert\|passed\|failed" gpurun_out/pytest_trace.log | head -30
timeout 300 python tools/trace_diff.py 128 0.6 2>&1 | grep -v Warn | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/trace_launches.csv python tools/trace_probe.py 64 256 1024 > gpurun_out/ncu_trace.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/trace_launches.csv')) if len(r) > 10 and r[0].isdigit()]
# split into forwards by the trace_grid_march kernel
idx = [i for i, r in enumerate(rows) if 'trace_grid_march' in r[4]]
import collections
for n, i in enumerate(idx):
    if n % 8 != 7: continue      # last forward of each resolution (3 warm + 5 timed)
    seg = rows[i - 1:(idx[n + 1] - 1 if n + 1 < len(idx) else len(rows))]
    agg = collections.OrderedDict()
    for r in seg:
        k = r[4].split('(')[0][-40:]
        t = float(r[-1]) / 1e3
        big = t > 10
        key = k + (' [ran]' if big else ' [idle]') if 'mlp_tc' in k else k
        a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += t
    print('--- forward', n, 'sum us', round(sum(v[1] for v in agg.values()), 1))
    for k, v in agg.items(): print(f'   {k:60s} x{v[0]:3d} {v[1]:9.1f} us')
PY
timeout 300 python tools/trace_probe.py 64 128 256 512 1024 2>&1 | grep trace
