python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --cpu-baseline 0 --steps 10 --warmup 3 > gpurun_out/r_A.json 2> gpurun_out/r_A.err
python -c "
import json
d=json.load(open('gpurun_out/r_A.json')); print(d['ms_per_step'], d['phases_ms_per_iteration'], d['e2e']['value'])
" || tail -5 gpurun_out/r_A.err
python tools/trace_factor.py final 2>&1 | tail -34
