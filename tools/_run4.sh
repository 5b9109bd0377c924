mkdir -p gpurun_out/s2
python -m pytest tests -m gpu -x -q -k "spa" > gpurun_out/s2/pytest_spa.log 2>&1; tail -2 gpurun_out/s2/pytest_spa.log
python tools/quick_bench.py 16384 jpl.4096.4.5 spa
python tools/quick_bench.py 32768 jpl.1024.4.5 spa
