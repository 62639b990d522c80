for cfg in "A:" "B:SFX_KC_FIXED=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs python bench.py --cpu-baseline 0 --steps 10 --warmup 3 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg', d['ms_per_step'], d['phases_ms_per_iteration'], d['e2e']['value'])"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
