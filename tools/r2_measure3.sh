#!/bin/bash
# ncu --set full of the Viterbi / Forward / domain kernels of config 3 (narrow and wide classes), summarised on the box
set -u
O=gpurun_out
mkdir -p $O
timeout 500 ncu --profile-from-start off --set full --clock-control none -k 'regex:vit16_wide_kernel<20>|fwd_wide_kernel|domain_kernel|vit16_kernel<24>|fwd_kernel_t<48>|vit16_wide_kernel<24>' -c 6 -o /tmp/r2_c3b python tools/run_config.py --config 3 --fraction 0.0131072 --profile-region > $O/r2_ncu_c3b.log 2>&1
python tools/ncu_summary.py /tmp/r2_c3b.ncu-rep > $O/r2_ncu_config3_vit_fwd_domain.md 2>&1
grep -A 12 "^## " $O/r2_ncu_config3_vit_fwd_domain.md | grep "^## \|time_duration\|registers\|issue_active\|pipe_alu\|warps_active" | cut -c1-150
du -sh $O
