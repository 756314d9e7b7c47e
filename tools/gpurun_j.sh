#!/bin/bash
# round-2 call J (1 GPU): ncu evidence -- launch list of a bench run, full captures of every loop kernel
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --no-e2e > gpurun_out/r2j_launches_bench.log 2>&1
tail -2 gpurun_out/r2j_launches_bench.log | cut -c1-300
# fused CG loop: calc_pw + calc_ur (skip the first iterations: cold)
timeout 600 $NCU -k regex:k_cg_calc -s 24 -c 4 -o gpurun_out/r2j_cg_fused -f python tools/profile_cg.py --iters 20 > gpurun_out/r2j_ncu.log 2>&1
# three-kernel loop: calc_w, calc_ur, calc_p
timeout 600 $NCU -k regex:k_cg_calc -s 30 -c 3 -o gpurun_out/r2j_cg_three -f python tools/profile_cg.py --iters 20 --fused 0 >> gpurun_out/r2j_ncu.log 2>&1
# one-pass Chebyshev / PPCG kernels
timeout 600 $NCU -k regex:k_fused_stencil -s 4 -c 2 -o gpurun_out/r2j_cheby -f python tools/profile_cg.py --iters 60 --solver cheby >> gpurun_out/r2j_ncu.log 2>&1
timeout 600 $NCU -k regex:k_fused_stencil -s 4 -c 2 -o gpurun_out/r2j_ppcg -f python tools/profile_cg.py --iters 40 --solver ppcg >> gpurun_out/r2j_ncu.log 2>&1
# the vectorised skeleton (field_summary, 2norm, copy, finalise) and the one-pass cg_init
timeout 600 $NCU -k regex:"k_vec|k_generic" -c 24 -o gpurun_out/r2j_setup -f python tools/profile_cg.py --iters 2 >> gpurun_out/r2j_ncu.log 2>&1
grep -c "==PROF==" gpurun_out/r2j_ncu.log; tail -3 gpurun_out/r2j_ncu.log
ls -la gpurun_out/*.ncu-rep
