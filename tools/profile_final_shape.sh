# Round-2 evidence run (on the GPU box, under gpurun): launch list + ncu --set full captures of the hot kernels of
# `python bench.py` at Final-shape, then the bench lines themselves.  Outputs go to gpurun_out/; summarise them with
#   python tools/ncu_summary.py gpurun_out/launches_r02.csv gpurun_out/prof_*_r02.ncu-rep > profiles/r02_final_shape_ncu_summary.txt
#   python tools/make_traffic_json.py r02
set -x
R=${R:-r02}
B="python bench.py --steps 2 --warmup 1 --cpu-baseline 0 --secondary 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv $B > gpurun_out/pp0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:large_factor -s 3 -c 1 -o gpurun_out/prof_factor_$R -f $B > gpurun_out/pp1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:schur_s9 -s 2 -c 1 -o gpurun_out/prof_s9_$R -f $B > gpurun_out/pp2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:linearize_bal -s 16 -c 1 -o gpurun_out/prof_lin_$R -f $B > gpurun_out/pp3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:schur_w_rhs -s 2 -c 1 -o gpurun_out/prof_w_$R -f $B > gpurun_out/pp4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bal_point_finalize -s 2 -c 1 -o gpurun_out/prof_fin_$R -f $B > gpurun_out/pp5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:schur_cinv -s 2 -c 1 -o gpurun_out/prof_cinv_$R -f $B > gpurun_out/pp6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:schur_back_accum -s 2 -c 1 -o gpurun_out/prof_back_$R -f $B > gpurun_out/pp7.log 2>&1
./tools/micro/l2_peak > gpurun_out/l2_peak_$R.json
python bench.py > gpurun_out/bench_default_$R.json 2> gpurun_out/bench_default_$R.err
ls -la gpurun_out/*$R*
