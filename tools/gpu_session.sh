#!/bin/bash
# One GPU-box session: smoke, parity tests, bench, kernel launch list.  Everything is logged
# under gpurun_out/ (merged back by gpurun).  Each stage has its own timeout.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest"; timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q --timeout 900 ${PYTEST_ARGS} > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest.log
if [ -z "$SKIP_BENCH" ]; then
echo "== bench"; timeout 900 python bench.py --steps ${BENCH_STEPS:-10} ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if [ -n "$RUN_NCU_LIST" ]; then
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu ${BENCH_ARGS} > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_bench.log
fi
if [ -n "$RUN_NCU_FULL" ]; then
echo "== ncu full capture of ${NCU_KERNEL:-mlp_tc_kernel}"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-mlp_tc_kernel} -s 2 -c 1 -f -o gpurun_out/prof_${NCU_KERNEL:-mlp_tc_kernel} python bench.py --steps 1 --warmup 1 --no-cpu ${BENCH_ARGS} > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
