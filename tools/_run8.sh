mkdir -p gpurun_out/s2
python tools/stream_probe.py 8192 2>&1 | tail -1
python tools/stream_probe.py 32768 1920.1280.3.303 min 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q -k "stream or golden or trace or edge or cached" 2>&1 | tail -2
