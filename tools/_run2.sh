mkdir -p gpurun_out/s2
python -m pytest tests -m gpu -x -q -k "resident" > gpurun_out/s2/pytest_res.log 2>&1; tail -2 gpurun_out/s2/pytest_res.log
for so in "" build/libldpc_b200_np4.so; do
  for code in jpl.4096.4.5 jpl.1024.4.5; do
    LDPC_B200_SO=$so python tools/quick_bench.py 32768 $code min
  done
  LDPC_B200_SO=$so python tools/quick_bench.py 16384 jpl.4096.4.5 spa
done 2>&1 | tee gpurun_out/s2/quick2.log
