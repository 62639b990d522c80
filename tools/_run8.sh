for v in "" mb5 mb6; do
  if [ -n "$v" ]; then export SFX_LIB=$PWD/symforce_b200/lib/libsfx_$v.so; fi
  echo "== variant ${v:-default(4)}"
  python tools/time_linearize.py final 2>&1 | grep "skip= 0\|skip= 7" | head -3
  python bench.py --cpu-baseline 0 --steps 10 --warmup 3 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms_per_iteration'])"
done
