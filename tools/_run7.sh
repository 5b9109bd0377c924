mkdir -p gpurun_out/s2
ncu --set full --clock-control none --import-source on -k regex:k_cn_tma -s 3 -c 1 -o gpurun_out/s2/prof_stream_cn_tma -f python tools/stream_probe.py 8192 > gpurun_out/s2/ncu_stream_cn.log 2>&1
tail -3 gpurun_out/s2/ncu_stream_cn.log
