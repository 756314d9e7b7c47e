#!/bin/bash
# round-2 call S (1 GPU): ncu --set full of the one-pass Chebyshev / PPCG kernels (4000x4000: 833 CG pre-steps first)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:k_fused_stencil -s 20 -c 3 -o gpurun_out/r2s_cheby -f python tools/profile_cg.py --iters 2500 --solver cheby > gpurun_out/r2s_ncu.log 2>&1
timeout 400 $NCU -k regex:k_fused_stencil -s 20 -c 3 -o gpurun_out/r2s_ppcg -f python tools/profile_cg.py --iters 2000 --solver ppcg >> gpurun_out/r2s_ncu.log 2>&1
grep -c "==PROF==" gpurun_out/r2s_ncu.log; grep "iters" gpurun_out/r2s_ncu.log; ls -la gpurun_out/r2s_*.ncu-rep
