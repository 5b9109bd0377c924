mkdir -p gpurun_out/ev
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/stream_probe.py 300 1920.1280.3.303 min auto 2.0 > gpurun_out/ev/san_mem_onchip.log 2>&1; echo "memcheck onchip rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/stream_probe.py 300 jpl.4096.4.5 min stream 3.5 > gpurun_out/ev/san_mem_stream.log 2>&1; echo "memcheck stream rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/stream_probe.py 600 jpl.4096.4.5 spa auto 3.5 > gpurun_out/ev/san_mem_resident.log 2>&1; echo "memcheck resident spa rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/stream_probe.py 64 1920.1280.3.303 min auto 2.0 > gpurun_out/ev/san_race_onchip.log 2>&1; echo "racecheck onchip rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/stream_probe.py 64 jpl.1024.4.5 min auto 3.5 > gpurun_out/ev/san_race_resident.log 2>&1; echo "racecheck resident rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/stream_probe.py 128 moon.7.13 min stream 3.0 > gpurun_out/ev/san_race_stream.log 2>&1; echo "racecheck stream rc=$?"
tail -3 gpurun_out/ev/san_*.log
