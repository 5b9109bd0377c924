mkdir -p gpurun_out/s2
python tools/stream_probe.py 8192 2>&1 | tail -1
python tools/stream_probe.py 8192 jpl.4096.4.5 spa 2>&1 | tail -1
python tools/stream_probe.py 32768 1920.1280.3.303 min 2>&1 | tail -1
python tools/stream_probe.py 32768 1920.1280.3.303 spa 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/s2/stream_launches4.csv python tools/stream_probe.py 8192 > /dev/null 2>&1
