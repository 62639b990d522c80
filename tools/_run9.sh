python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "A:" "B:SFX_SOLVE_V1=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs python bench.py --cpu-baseline 0 --steps 10 --warmup 3 > gpurun_out/r_$name.json 2> gpurun_out/r_$name.err
  python -c "
import json
d=json.load(open('gpurun_out/r_$name.json')); print('$cfg', d['ms_per_step'], d['phases_ms_per_iteration'], d['e2e']['value'])
" || tail -5 gpurun_out/r_$name.err
done
