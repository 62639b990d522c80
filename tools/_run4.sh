python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "A:" "B:SFX_SCHUR_V2=1" "C:SFX_POINT_ATOMICS=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs python bench.py --cpu-baseline 0 --steps 10 --warmup 3 > gpurun_out/r_$name.json 2> gpurun_out/r_$name.err
  python -c "
import json
d=json.load(open('gpurun_out/r_$name.json')); print('$cfg', d['ms_per_step'], d['phases_ms_per_iteration'])
" || tail -5 gpurun_out/r_$name.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 1 --cpu-baseline 0 > gpurun_out/p_v3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:schur_s3 -s 2 -c 1 -o gpurun_out/prof_s3 -f python bench.py --steps 2 --warmup 1 --cpu-baseline 0 > gpurun_out/p_s3.log 2>&1
