# A/B of the variable-phase plans (build/libldpc_b200_*.so variants): decisions CRC + kernel time per variant
# usage: bash tools/ab_vn.sh [variant to run the GPU test suite with]
O=gpurun_out/ab; mkdir -p $O
for so in default $(ls build/libldpc_b200_*.so); do
  if [ $so = default ]; then unset LDPC_B200_SO; else export LDPC_B200_SO=$PWD/$so; fi
  case $so in
    *lat_*) python tools/latency_probe.py 2.0 2>&1 | grep "decode_device\|n_frames=1:" | sed "s|^|$so |";;
    *spa_*) python tools/quick_bench.py 16384 jpl.4096.4.5 spa;;
    default) python tools/quick_bench.py 32768 jpl.4096.4.5 min; python tools/quick_bench.py 16384 jpl.1024.4.5 min; python tools/quick_bench.py 16384 jpl.4096.4.5 spa
             python tools/latency_probe.py 2.0 2>&1 | grep "decode_device\|n_frames=1:" | sed "s|^|$so |";;
    *) python tools/quick_bench.py 32768 jpl.4096.4.5 min; python tools/quick_bench.py 16384 jpl.1024.4.5 min;;
  esac
done > $O/ab_vn.log 2>&1
cat $O/ab_vn.log
if [ -n "$1" ]; then
export LDPC_B200_SO=$PWD/build/libldpc_b200_$1.so
python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_tester_cli > $O/pytest_variant.log 2>&1; tail -5 $O/pytest_variant.log
fi
