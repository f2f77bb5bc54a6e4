mkdir -p gpurun_out
{
timeout 200 python -c "import torch; print('torch ok', torch.cuda.is_available())"
for v in "SDFR_TC_PAIR=1" "SDFR_TC_PAIR=0"; do
env $v PROBE_N=1 timeout 40 python tools/perf_probe.py 2>&1 | grep -v Warn | grep -E "^\[|rror|Trace|File" | grep coarse
echo "rc=$? ($v)"
done
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
} > gpurun_out/p1.log 2>&1
cat gpurun_out/p1.log
