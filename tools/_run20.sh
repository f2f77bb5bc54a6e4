SDFR_LIB=sdflabel_b200/libsdfr_dbg.so timeout 300 python tools/trace_diff.py 128 0.6 2>&1 | grep -v Warn | grep "ray 7340\|hits" | head -80
