mkdir -p gpurun_out/s2
python tools/one_decode.py spa 2.0 16384 1 > gpurun_out/s2/plain_spa_v2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_resident -s 1 -c 1 -o gpurun_out/s2/prof_spa_v2 -f python tools/one_decode.py spa 2.0 16384 1 > gpurun_out/s2/ncu_spa_v2.log 2>&1
cat gpurun_out/s2/plain_spa_v2.log
