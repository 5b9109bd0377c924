set -x
mkdir -p gpurun_out/s2
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s2/pytest_gpu.log 2>&1
tail -3 gpurun_out/s2/pytest_gpu.log
( time python bench.py ) > gpurun_out/s2/bench_full.json 2> gpurun_out/s2/bench_full.err
tail -c 600 gpurun_out/s2/bench_full.json
( time python bench.py --impl reference ) > gpurun_out/s2/bench_ref.json 2> gpurun_out/s2/bench_ref.err
tail -c 400 gpurun_out/s2/bench_ref.json
python tools/stream_probe.py 8192 > gpurun_out/s2/stream_probe.log 2>&1
python tools/stream_probe.py 8192 jpl.4096.4.5 spa >> gpurun_out/s2/stream_probe.log 2>&1
python tools/stream_probe.py 32768 1920.1280.3.303 min >> gpurun_out/s2/stream_probe.log 2>&1
cat gpurun_out/s2/stream_probe.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2/smoke.log 2>&1; tail -3 gpurun_out/s2/smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s2/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/s2/ncu_bench.log 2>&1
