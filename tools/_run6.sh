mkdir -p gpurun_out/s2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s2/pytest_gpu2.log 2>&1; tail -2 gpurun_out/s2/pytest_gpu2.log
python tools/stream_probe.py 8192 2>&1 | tee gpurun_out/s2/stream_probe2.log
python tools/stream_probe.py 8192 jpl.4096.4.5 spa 2>&1 | tee -a gpurun_out/s2/stream_probe2.log
python tools/stream_probe.py 32768 1920.1280.3.303 min 2>&1 | tee -a gpurun_out/s2/stream_probe2.log
python tools/stream_probe.py 32768 1920.1280.3.303 spa 2>&1 | tee -a gpurun_out/s2/stream_probe2.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/s2/stream_launches3.csv python tools/stream_probe.py 8192 > /dev/null 2>&1
