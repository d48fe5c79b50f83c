#!/bin/bash
# Round profile: launch list of a short bench (shares per kernel; ncu needs ~0.17 s per launch, keep -c small) + full captures of the
# dominant kernels (6 Lc launches then 6 L pattern launches in tools/ncu_mma.py: -s skips, -c 1 captures one).
set -x
R=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${R}_launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
python tools/launch_shares.py gpurun_out/${R}_launches_c2.csv "ncu launch list, bench.py --workload c2 --steps 1 --warmup 1 (first 3000 launches), ${R} build" > gpurun_out/${R}_launch_shares_c2.txt
ncu --set full --clock-control none --import-source on -k regex:bsr_spmm_mma_native -s 3 -c 1 -o gpurun_out/${R}_spmm_mma_native_Lc_b64 python tools/ncu_mma.py torus 1000000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bsr_spmm_mma_native -s 9 -c 1 -o gpurun_out/${R}_spmm_mma_pattern_L_b64 python tools/ncu_mma.py torus 1000000 > /dev/null 2>&1
for f in ${R}_spmm_mma_native_Lc_b64 ${R}_spmm_mma_pattern_L_b64; do python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/${f}_summary.txt 2>&1; done
ls -la gpurun_out | tail -12
