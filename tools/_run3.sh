mkdir -p gpurun_out/s2
python tools/one_decode.py min 2.0 32768 1 > gpurun_out/s2/plain_v6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_resident -s 1 -c 1 -o gpurun_out/s2/prof_resident_v6 -f python tools/one_decode.py min 2.0 32768 1 > gpurun_out/s2/ncu_v6.log 2>&1
tail -2 gpurun_out/s2/plain_v6.log gpurun_out/s2/ncu_v6.log
