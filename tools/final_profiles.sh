#!/bin/bash
# Round-end measurement pass on one B200: bench (both arms), ncu launch list of the same command, full captures of the two top kernels.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/final_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"schur_mma_kernel" -s 2 -c 1 -o gpurun_out/final_schur_mma -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/final_ncu_schur.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spmv_tma_kernel" -s 1 -c 1 -o gpurun_out/final_spmv -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/final_ncu_spmv.log 2>&1
ncu --set full --clock-control none -k regex:"coeff_w_kernel|schur_pairs_kernel|build_pl_kernel|backsub_accum_kernel" -s 4 -c 4 -o gpurun_out/final_others -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/final_ncu_others.log 2>&1
ls -la gpurun_out/final_*
