python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "A:" "B:SFX_KC=1" "C:SFX_KC=6" "D:SFX_KC=8" "E:SFX_SCHUR_V2=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs python bench.py --cpu-baseline 0 --steps 10 --warmup 3 > gpurun_out/r_$name.json 2> gpurun_out/r_$name.err
  python -c "
import json
d=json.load(open('gpurun_out/r_$name.json')); print('$cfg', d['ms_per_step'], d['phases_ms_per_iteration'])
" || tail -5 gpurun_out/r_$name.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 2 --warmup 1 --cpu-baseline 0 > gpurun_out/p_v4.log 2>&1
