mkdir -p gpurun_out/s2
ncu --set full --clock-control none --import-source on -k regex:k_onchip -s 1 -c 1 -o gpurun_out/s2/prof_onchip -f python tools/stream_probe.py 65536 1920.1280.3.303 min auto > gpurun_out/s2/ncu_onchip.log 2>&1
tail -2 gpurun_out/s2/ncu_onchip.log
