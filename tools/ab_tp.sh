# A/B of the tester pattern (T threads x one frame per call) between the shipped library and build/libldpc_b200_lat_*.so
O=gpurun_out/ab; mkdir -p $O
for r in 1 2 3; do
for so in default $(ls build/libldpc_b200_lat_*.so); do
  if [ $so = default ]; then unset LDPC_B200_SO; else export LDPC_B200_SO=$PWD/$so; fi
  python tools/tester_pattern_probe.py ldpc/b200-minsum/jpl.4096.4.5/20/4/5 1,16,32 | sed "s|^|$so |"
done; done > $O/ab_tp.log 2>&1
cat $O/ab_tp.log
