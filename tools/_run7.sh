mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trace.py -m gpu -q --timeout 300 2>&1 | tail -8 | grep -v Warn
timeout 300 python tools/_prof_small.py 2>&1 | grep -v Warn
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --quick --frames 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
for k in mlp_tc_band_pair_kernel mlp_tc_coarse_pair_kernel splat_forward_kernel select_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o /tmp/prof_$k python bench.py --steps 1 --warmup 1 --no-cpu --quick --frames 0 > gpurun_out/ncu_full_$k.log 2>&1; echo "ncu full $k rc=$?"
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/r02_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$k.ncu-rep --page details > gpurun_out/r02_${k}_details.txt 2>/dev/null
done
ls -la gpurun_out | head -30
