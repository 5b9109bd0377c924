# A/B of the latency instances behind the mailbox: one-frame calls from one thread, three repetitions per library
O=gpurun_out/ab; mkdir -p $O
for so in default $(ls build/libldpc_b200_lat_*.so); do
  if [ $so = default ]; then unset LDPC_B200_SO; else export LDPC_B200_SO=$PWD/$so; fi
  for r in 1 2 3; do python tools/one_mailbox.py min 2.0 3000 | sed "s|^|$so |"; done
  python tools/one_mailbox.py min 4.0 3000 | sed "s|^|$so |"
  python tools/one_mailbox.py spa 2.0 2000 | sed "s|^|$so |"
done > $O/ab_lat.log 2>&1
cat $O/ab_lat.log
