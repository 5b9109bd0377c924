N=$1
mkdir -p gpurun_out/ev
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-extra > gpurun_out/ev/bench_${N}gpu.json 2> gpurun_out/ev/bench_${N}gpu.err
tail -c 1200 gpurun_out/ev/bench_${N}gpu.json; tail -3 gpurun_out/ev/bench_${N}gpu.err
