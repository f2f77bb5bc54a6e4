timeout 300 ncu --set full --clock-control none -k regex:splat_forward_kernel -s 2 -c 1 -f -o /tmp/prof_s32 python tools/_one_small.py 32 > gpurun_out/ncu_s32.log 2>&1; echo rc=$?
ncu -i /tmp/prof_s32.ncu-rep --page details 2>/dev/null | grep -E "Duration|Executed Ipc|Issue Slots|Warp Cycles Per Issued|Grid Size|Registers|Stall|stall|Eligible|Branch|Divergent|Avg\. Active Threads|Instructions|Local|Theoretical Occ|Achieved Occ|L1/TEX Hit|Shared" | head -40
ncu -i /tmp/prof_s32.ncu-rep --page details 2>/dev/null | grep -B2 -A12 "Warp State Statistics" | head -50
