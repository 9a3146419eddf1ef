#!/bin/bash
# ncu --set full of the Forward kernels of config 3: fwd_kernel_t<48> (shared-memory rows, compile-time B) and the new fwd_wide_kernel
set -u
O=gpurun_out
mkdir -p $O
timeout 120 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k 'regex:fwd_kernel_t<.int.48>|fwd_wide_kernel' -c 4 -o /tmp/r2_fw python tools/run_config.py --config 3 --fraction 0.0032768 --profile-region > $O/r2_ncu_fw.log 2>&1
python tools/ncu_summary.py /tmp/r2_fw.ncu-rep > $O/r2_ncu_config3_fwd.md 2>&1
grep "^## \|time_duration\|registers\|issue_active\|warps_active\|dram__bytes" $O/r2_ncu_config3_fwd.md | cut -c1-150
