#!/bin/bash
# parity of the shared-memory fwd_wide_kernel + its share of a config-3 chunk (compare profiles/r2_config3_launch_shares.txt)
set -u
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_wide.py tests/test_gpu_configs.py -q -m gpu -x 2>&1 | tail -25 | cut -c1-220
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2b_config3_launches.csv python tools/run_config.py --config 3 --fraction 0.0131072 --profile-region > $O/r2b_c3.log 2>&1
python tools/launch_shares.py $O/r2b_config3_launches.csv > $O/r2b_config3_launch_shares.txt; head -12 $O/r2b_config3_launch_shares.txt
tail -1 $O/r2b_c3.log | cut -c1-900
