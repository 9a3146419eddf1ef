#!/bin/bash
# Round-2 measurement batch (one gpurun call, one GPU): tests, per-config parity, bench lines, ncu launch list + full captures,
# full-size config runs.  Everything lands in gpurun_out/ (copied to profiles/ by hand afterwards).
set -u
O=gpurun_out
mkdir -p $O
step() { echo "=== $1 ($(date +%T))"; }
step "gpu tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 > $O/r2_gpu_tests.log; tail -2 $O/r2_gpu_tests.log
step "config parity"; timeout 900 python -m tests.config_parity --config 2 3 4 5 --msv-impl --out $O/r2_config_parity.json > /dev/null 2> $O/r2_config_parity.log; tail -3 $O/r2_config_parity.log
step "bench"; timeout 600 python bench.py > $O/r2_bench.json 2> $O/r2_bench.err; cut -c1-400 $O/r2_bench.json
step "bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference.json 2>/dev/null; cut -c1-300 $O/r2_bench_reference.json
step "ncu launch list (bench)"; timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-region > $O/r2_ncu_bench.log 2>&1; python tools/launch_shares.py $O/r2_launches.csv > $O/r2_launch_shares.txt; head -12 $O/r2_launch_shares.txt
step "ncu full: msv_pack_kernel (bench step)"; timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:msv_pack_kernel -c 1 -o $O/r2_msv_pack python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-region > $O/r2_ncu_pack.log 2>&1; ls -la $O/r2_msv_pack.ncu-rep
step "ncu full: config-3 kernels"; timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:msv_wide_kernel|vit16_wide_kernel|fwd_wide_kernel|domain_kernel|msv_exact_wide' -c 12 -o $O/r2_config3_kernels python tools/run_config.py --config 3 --fraction 0.0131072 --profile-region > $O/r2_ncu_c3.log 2>&1; ls -la $O/r2_config3_kernels.ncu-rep
step "config 2 full"; timeout 600 python tools/run_config.py --config 2 --out $O/r2_config2_full.json 2>/dev/null | cut -c1-300
step "config 4 full"; timeout 600 python tools/run_config.py --config 4 --out $O/r2_config4_full.json 2>/dev/null | cut -c1-300
step "config 3 full"; timeout 900 python tools/run_config.py --config 3 --out $O/r2_config3_full.json 2>/dev/null | cut -c1-300
step "done"
