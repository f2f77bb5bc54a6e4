mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trace.py -m gpu -q --timeout 300 > gpurun_out/pytest_trace.log 2>&1; echo "pytest rc=$?"; grep -n "^E  *Assert\|^E  *assert\|passed\|failed" gpurun_out/pytest_trace.log | head -30
timeout 300 python tools/trace_probe.py 64 128 256 512 1024 2>&1 | grep trace
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --frames 512 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
f=d['frames']; print('frames', {k:f[k] for k in f if k!='workload'})
PY
