N=$1
SFX_E2E_DIAG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --cpu-baseline 0 2> gpurun_out/e2e_diag_$N.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('${N}gpu', d['ms_per_step'], d['phases_ms_per_iteration'], d['e2e'])"
grep "e2e rank" gpurun_out/e2e_diag_$N.err | sort | head -8
nproc; lscpu | grep -i "numa\|socket\|model name" | head -6
