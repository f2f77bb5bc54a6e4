mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --frames 512 --no-cpu --sustain 0.5 > gpurun_out/bench_same_n1.json 2> gpurun_out/bench_same_n1.err; echo "bench n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --frames 512 --no-cpu --sustain 0.5 > gpurun_out/bench_same_n2.json 2> gpurun_out/bench_same_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_same_n1.json', 'gpurun_out/bench_same_n2.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    fr=d['frames']; print(f, {k:fr[k] for k in fr if k not in ('workload',)})
PY
