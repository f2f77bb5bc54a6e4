PROBE_N=1850,37,3700,7400,58000,58016 timeout 400 python tools/band_probe.py 2>&1 | grep -v Warn | tail -22
