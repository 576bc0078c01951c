#!/bin/bash
# Round-2 profile captures (run under gpurun on ONE GPU; numbers printed under ncu are never bench values).
# Launch lists (per-launch device time, cold-cache and serialised: compare SHARES) and full captures of the top kernels,
# exported to CSV on the box (the .ncu-rep files stay there: gpurun_out/ is capped at 64 MiB).
set -x
O=gpurun_out
NCU="ncu --clock-control none"
B="--no-flux --no-sd3 --no-qwen --no-sparse --no-fp8-attention --no-torch-baseline --no-cpu-baseline --steps 1 --warmup 3"
# one timed step each: bench.py brackets it with cudaProfilerStart/Stop when FDM_BENCH_PROFILE=1
FDM_BENCH_PROFILE=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r02_wan_step_launches.csv python bench.py $B > $O/ncu_wan_stdout.log 2>&1
# FLUX step, launched eagerly (the bench replays it as a CUDA graph)
FDM_BENCH_PROFILE=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r02_flux_step_launches.csv python bench.py --workload flux --no-graph $B > $O/ncu_flux_stdout.log 2>&1
# top kernels, full metric set, raw pages only (the tools launch each kernel twice: the second launch is captured)
$NCU --set full --import-source on -k regex:attn_fwd -s 1 -c 1 -o /tmp/p_attn_wan python tools/attn_one.py wan > $O/ncu_attn_wan.log 2>&1
ncu -i /tmp/p_attn_wan.ncu-rep --page raw --csv > $O/r02_attn_wan_raw.csv
$NCU --set full --import-source on -k regex:attn_fwd -s 1 -c 1 -o /tmp/p_attn_flux python tools/attn_one.py flux > $O/ncu_attn_flux.log 2>&1
ncu -i /tmp/p_attn_flux.ncu-rep --page raw --csv > $O/r02_attn_flux_raw.csv
$NCU --set full -k regex:gemm_w8a8 -s 1 -c 1 -o /tmp/p_gemm python tools/gemm_one.py > $O/ncu_gemm.log 2>&1
ncu -i /tmp/p_gemm.ncu-rep --page raw --csv > $O/r02_gemm_raw.csv
$NCU --set full -k regex:"quant|rmsnorm|rope|gelu" -c 12 -o /tmp/p_elem python tools/elementwise_one.py > $O/ncu_elem.log 2>&1
ncu -i /tmp/p_elem.ncu-rep --page raw --csv > $O/r02_elem_raw.csv
$NCU --set full -k regex:"quant" -s 4 -c 4 -o /tmp/p_lnq python tools/lnq_one.py > $O/ncu_lnq.log 2>&1
ncu -i /tmp/p_lnq.ncu-rep --page raw --csv > $O/r02_lnq_raw.csv
ls -la $O | tail -12
