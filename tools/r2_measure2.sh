#!/bin/bash
# Round-2 measurement batch, compact (ncu reports are summarised on the box and deleted: gpurun_out must stay small)
set -u
O=gpurun_out
mkdir -p $O
step() { echo "=== $1 ($(date +%T))"; }
step "new gpu tests"; timeout 300 python -m pytest tests/test_gpu_config1.py tests/test_gpu_merge.py -q -m gpu -k "uniquified or merge" 2>&1 | tail -3
step "config parity"; timeout 600 python -m tests.config_parity --config 2 3 4 5 --msv-impl --out $O/r2_config_parity.json > /dev/null 2> $O/r2_config_parity.log; tail -2 $O/r2_config_parity.log | cut -c1-300
step "bench"; timeout 400 python bench.py > $O/r2_bench.json 2> $O/r2_bench.err; cut -c1-200 $O/r2_bench.json
step "bench reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference.json 2>/dev/null; cut -c1-200 $O/r2_bench_reference.json
step "ncu launch list (bench)"; timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-region > $O/r2_ncu_bench.log 2>&1; python tools/launch_shares.py $O/r2_launches.csv > $O/r2_launch_shares.txt; head -4 $O/r2_launch_shares.txt
step "ncu full: msv_pack_kernel (32,768-read step)"; timeout 400 ncu --profile-from-start off --set full --clock-control none -k regex:msv_pack_kernel -c 1 -o /tmp/r2_pack python bench.py --steps 1 --warmup 3 --reads-per-step 32768 --no-cpu-baseline --profile-region > $O/r2_ncu_pack.log 2>&1; python tools/ncu_summary.py /tmp/r2_pack.ncu-rep > $O/r2_ncu_msv_pack.md 2>&1; head -8 $O/r2_ncu_msv_pack.md
step "ncu full: config-3 wide kernels"; timeout 600 ncu --profile-from-start off --set full --clock-control none -k 'regex:msv_wide_kernel|vit16_wide_kernel|fwd_wide_kernel|domain_kernel|msv_exact_wide' -c 7 -o /tmp/r2_c3 python tools/run_config.py --config 3 --fraction 0.0131072 --profile-region > $O/r2_ncu_c3.log 2>&1; python tools/ncu_summary.py /tmp/r2_c3.ncu-rep > $O/r2_ncu_config3_kernels.md 2>&1; grep "^## " $O/r2_ncu_config3_kernels.md | cut -c1-120
du -sh $O
step "done"
