#!/bin/bash
# Round-2 measurement pass on one B200: the full bench line (with CPU baseline and parity block), then -- each only after
# its own command has exited 0 without ncu -- the ncu launch list of the bench command and `ncu --set full` captures of the
# two dominant kernels (1D: newton1d_kernel on a saturated launch; 3D: gmres_cluster_kernel at batch 128).
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bench_n1.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config1 > gpurun_out/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config1 > gpurun_out/r02_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_1d.py --mesh 1e-6 --voltages 1280 --reps 2 > gpurun_out/r02_prof1d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:newton1d -s 1 -c 1 -f -o gpurun_out/r02_newton1d \
    python tools/prof_1d.py --mesh 1e-6 --voltages 1280 --reps 2 > gpurun_out/r02_ncu_newton1d.log 2>&1
echo "ncu 1d rc=$?"; cat gpurun_out/r02_prof1d.log
python tools/prof_3d.py --batch 128 --steps 1 > gpurun_out/r02_prof3d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gmres_cluster -s 1 -c 1 -f -o gpurun_out/r02_gmres \
    python tools/prof_3d.py --batch 128 --steps 1 > gpurun_out/r02_ncu_gmres.log 2>&1
echo "ncu 3d rc=$?"; cat gpurun_out/r02_prof3d.log
ls -la gpurun_out/*.ncu-rep
# batch-lane 3D assembly: the eight launches of one J+F assembly at batch 128 (profiles/r02_asm_lanes_ncu_summary.md)
python tools/prof_3d.py --batch 128 --steps 0 > gpurun_out/r02_prof3d_asm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"lanes|residual_gather" -s 8 -c 8 -f -o gpurun_out/r02_asm_lanes_final \
    python tools/prof_3d.py --batch 128 --steps 0 > gpurun_out/r02_ncu_asm.log 2>&1
echo "ncu asm rc=$?"; cat gpurun_out/r02_prof3d_asm.log
