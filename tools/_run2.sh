ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_v2.csv python bench.py --steps 2 --warmup 1 --cpu-baseline 0 > gpurun_out/p_v2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:schur_s2 -s 2 -c 1 -o gpurun_out/prof_s2 -f python bench.py --steps 2 --warmup 1 --cpu-baseline 0 > gpurun_out/p_s2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:schur_w_rhs -s 2 -c 1 -o gpurun_out/prof_w -f python bench.py --steps 2 --warmup 1 --cpu-baseline 0 > gpurun_out/p_w.log 2>&1
ls -la gpurun_out | tail -5
