mkdir -p gpurun_out
timeout 300 python tools/trace_probe.py 64 256 > gpurun_out/trace_probe.log 2>&1; cat gpurun_out/trace_probe.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/trace_launches.csv python tools/trace_probe.py 256 > gpurun_out/ncu_trace.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/trace_launches.csv')) if len(r) > 10 and r[0].isdigit()]
# last forward: take the final 200 launches
agg = collections.OrderedDict()
for r in rows[-120:]:
    print(r[4][:50], r[-1])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --frames 512 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
f=d['frames']; print('frames', {k:f[k] for k in f if k!='workload'})
PY
