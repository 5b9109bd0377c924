# One-GPU evidence run of round 2: bench (both arms), ncu launch lists + full captures.  Outputs under gpurun_out/ev2/.
# Two parts (gpurun brings back at most 64 MiB per call): `evidence_r02.sh A` = bench + RESIDENT captures, `B` = the rest.
set -x
O=gpurun_out/ev2; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
PART=${1:-AB}
if [[ $PART == *A* ]]; then
( time python bench.py --steps 20 --warmup 5 ) > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -4 $O/bench_1gpu.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > $O/plain_bench_short.json 2> $O/plain_bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > $O/ncu_launches.log 2>&1
$NCU -k regex:k_resident -s 2 -c 1 -o $O/prof_resident_min -f python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > $O/ncu_full_min.log 2>&1
python tools/one_decode.py spa 2.0 16384 1 > $O/plain_spa.log 2>&1 && \
$NCU -k regex:k_resident -s 1 -c 1 -o $O/prof_resident_spa -f python tools/one_decode.py spa 2.0 16384 1 > $O/ncu_full_spa.log 2>&1
python tools/quick_bench.py 16384 jpl.1024.4.5 min > $O/plain_min1024.log 2>&1 && \
$NCU -k regex:k_resident -s 3 -c 1 -o $O/prof_resident_min1024 -f python tools/quick_bench.py 16384 jpl.1024.4.5 min > $O/ncu_full_min1024.log 2>&1
python -c "from ecc_ldpc_b200 import buildinfo; print(buildinfo.kernel_source_sha256())" > $O/kernel_sources.sha256
fi
if [[ $PART == *B* ]]; then
python tools/one_mailbox.py min 2.0 200 > $O/plain_mailbox.log 2>&1 && \
$NCU -k regex:k_resident -s 4 -c 1 -o $O/prof_latency_min -f python tools/one_mailbox.py min 2.0 12 > $O/ncu_full_latency.log 2>&1
python tools/stream_probe.py 8192 > $O/plain_stream.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 100 --csv --log-file $O/launches_stream.csv python tools/stream_probe.py 8192 > /dev/null 2>&1
$NCU -k regex:k_cn_tma -s 3 -c 1 -o $O/prof_stream_cn -f python tools/stream_probe.py 8192 > $O/ncu_full_cn.log 2>&1
$NCU -k regex:k_vn_tma -s 3 -c 1 -o $O/prof_stream_vn -f python tools/stream_probe.py 8192 > $O/ncu_full_vn.log 2>&1
python tools/stream_probe.py 65536 1920.1280.3.303 min auto > $O/plain_onchip.log 2>&1 && \
$NCU -k regex:k_onchip -s 1 -c 1 -o $O/prof_onchip -f python tools/stream_probe.py 65536 1920.1280.3.303 min auto > $O/ncu_full_onchip.log 2>&1
python tools/latency_probe.py 2.0 > $O/latency_probe.log 2>&1
fi
ls -la $O
