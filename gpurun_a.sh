cd /root/repo
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/kernel_roofline.py 4000 > gpurun_out/kernels_roofline.txt 2>&1; tail -6 gpurun_out/kernels_roofline.txt
