python tools/stream_probe.py 65536 1920.1280.3.303 min auto | tail -1
python tools/stream_probe.py 65536 1920.1280.3.303 spa auto | tail -1
python tools/stream_probe.py 65536 1920.1280.3.303 min auto 2.5 | tail -1
python tools/stream_probe.py 262144 moon.7.13 min auto | tail -1
python tools/stream_probe.py 262144 moon.7.13 spa auto 3.0 | tail -1
python tools/stream_probe.py 16384 jpl.1024.4.5 min onchip | tail -1
timeout 900 python -m pytest tests -m gpu -x -q -k "onchip or bad_arg or edge" 2>&1 | tail -3
