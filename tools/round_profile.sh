#!/bin/bash
# Round profile: launch list of a short bench (shares per kernel) + full captures of the dominant kernels.
set -x
R=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${R}_launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bsr_spmm_mma_native -s 3 -c 1 -o gpurun_out/${R}_spmm_mma_native_Lc_b64 python tools/ncu_mma.py torus 1000000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bsr_spmm_v2 -s 3 -c 1 -o gpurun_out/${R}_spmm_v2_Lc_b64 python tools/ncu_one.py torus 1000000 Lc 64 0 > /dev/null 2>&1
ls -la gpurun_out
