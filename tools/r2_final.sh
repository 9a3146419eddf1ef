#!/bin/bash
# final round-2 check: whole GPU suite, smoke(), then ncu --set full of the wide-class Viterbi / Forward kernels of config 3
set -u
O=gpurun_out
mkdir -p $O
timeout 330 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | cut -c1-200 | tee $O/r2_final_gpu_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 110 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k 'regex:vit16_kernel<24>|fwd_kernel_t<48>|vit16_wide_kernel<20>' -c 3 -o /tmp/r2_vf python tools/run_config.py --config 3 --fraction 0.0032768 --profile-region > $O/r2_ncu_vf.log 2>&1
python tools/ncu_summary.py /tmp/r2_vf.ncu-rep > $O/r2_ncu_config3_vit16_fwd.md 2>&1
grep "^## \|time_duration\|registers\|issue_active\|pipe_alu\|warps_active" $O/r2_ncu_config3_vit16_fwd.md | cut -c1-150
