set -x
B="python bench.py --steps 2 --warmup 1 --cpu-baseline 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv $B > gpurun_out/pp0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:large_factor -s 6 -c 1 -o gpurun_out/prof_factor_r1d -f $B > gpurun_out/pp1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:schur_s9 -s 2 -c 1 -o gpurun_out/prof_s9_r1d -f $B > gpurun_out/pp2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:linearize_bal -s 2 -c 1 -o gpurun_out/prof_lin_r1d -f $B > gpurun_out/pp3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:schur_w_rhs -s 2 -c 1 -o gpurun_out/prof_w_r1d -f $B > gpurun_out/pp4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:large_factor -s 9 -c 1 -o gpurun_out/prof_factor_root_r1d -f $B > gpurun_out/pp5.log 2>&1
python bench.py > gpurun_out/bench_default_r1d.json 2> gpurun_out/bench_default_r1d.err
python bench.py --workload ladybug --cpu-baseline 0 > gpurun_out/bench_ladybug_r1d.json 2>> gpurun_out/bench_default_r1d.err
ls -la gpurun_out/*r1d*
