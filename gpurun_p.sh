cd /root/repo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01.csv python tools/profile_cg.py --iters 20 > gpurun_out/prof_launch.log 2>&1; tail -1 gpurun_out/prof_launch.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_cg_calc -s 24 -c 4 -f -o gpurun_out/prof_cg_r01 python tools/profile_cg.py --iters 20 > gpurun_out/prof_full.log 2>&1; tail -1 gpurun_out/prof_full.log
