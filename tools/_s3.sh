mkdir -p gpurun_out
{
for v in "SDFR_TC_CLUSTER_SMALL=1" "SDFR_TC_CLUSTER_SMALL=2" "SDFR_TC_CLUSTER_SMALL=4" "SDFR_TC_SMALL_NP=32 SDFR_TC_CLUSTER_SMALL=1" "SDFR_TC_SMALL_NP=32 SDFR_TC_CLUSTER_SMALL=2" "SDFR_TC_SMALL_NP=32 SDFR_TC_CLUSTER_SMALL=4" "SDFR_TC_SMALL_NP=64"; do
  env $v timeout 40 python tools/perf_probe.py 2>&1 | grep -v Warn | grep "^\["
done
} > gpurun_out/s3.log 2>&1
cat gpurun_out/s3.log
