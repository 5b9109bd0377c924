timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench.py 32768 jpl.4096.4.5 min
python tools/quick_bench.py 32768 jpl.1024.4.5 min
