# Multi-GPU evidence of round 2 (run under gpurun --gpus 8): bench.py at N = 2, 4, 8 with extras (H2D ceiling, simulate_multi leg),
# and the GPU suite on a multi-GPU box (two-device test, simulate_multi at n_gpus > 1).
O=gpurun_out/ev2; mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "N=1 rc=$?" >> $O/scale.log
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
  echo "N=$N rc=$?" >> $O/scale.log
done
python -m pytest tests -m gpu -q -k "two_devices or simulate_multi or shutdown" > $O/pytest_multi_gpu.log 2>&1; echo "pytest rc=$?" >> $O/scale.log
nvidia-smi topo -m > $O/topo.txt 2>&1; lscpu | head -25 > $O/lscpu.txt; numactl -H >> $O/lscpu.txt 2>&1
