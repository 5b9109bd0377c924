# Device-timed scaling only (no extras, no CPU leg): bench.py at N = 1, 2, 4, 8 on one 8-GPU box, ~15 s per N
O=gpurun_out/ev2; mkdir -p $O
python bench.py --steps 20 --warmup 5 --no-extra --no-cpu > $O/lean_1gpu.json 2> $O/lean_1gpu.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29620+N)) bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu > $O/lean_${N}gpu.json 2> $O/lean_${N}gpu.err
done
for N in 1 2 4 8; do python -c "
import json,sys
d=json.loads(open('$O/lean_${N}gpu.json').read().strip().splitlines()[-1]); print($N, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done
