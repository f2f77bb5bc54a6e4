mkdir -p gpurun_out
{
for v in "SDFR_TC_WIDE_DBG=16" "SDFR_TC_WIDE_DBG=0"; do
  env $v timeout 25 python tools/perf_probe.py 2>&1 | grep -v Warn | grep "^\[" | head -1
done
} > gpurun_out/s2.log 2>&1
tail -60 gpurun_out/s2.log
