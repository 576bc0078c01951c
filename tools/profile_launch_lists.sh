#!/bin/bash
O=gpurun_out
NCU="ncu --clock-control none"
B="--no-flux --no-sd3 --no-qwen --no-sparse --no-fp8-attention --no-torch-baseline --no-cpu-baseline --steps 1 --warmup 3"
FDM_BENCH_PROFILE=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r02_wan_step_launches.csv python bench.py $B > $O/ncu_wan_stdout.log 2>&1
FDM_BENCH_PROFILE=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r02_flux_step_launches.csv python bench.py --workload flux --no-graph $B > $O/ncu_flux_stdout.log 2>&1
ls -la $O/r02_*_launches.csv
