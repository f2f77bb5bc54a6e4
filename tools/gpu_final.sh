#!/bin/bash
# Round-end GPU session: parity suite, smoke, the bench line, ncu launch lists (refine step, trace mode).
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
echo "== pytest"; ( time timeout 1200 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench"; ( time timeout 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -5 gpurun_out/bench.err
echo "== bench reference arm"; ( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -4 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
echo "== ncu launch list: refine step"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --quick > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu launch list: trace 256"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/trace_launches.csv python tools/trace_probe.py 256 > gpurun_out/trace_ncu.log 2>&1; echo "ncu rc=$?"
