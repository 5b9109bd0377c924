mkdir -p gpurun_out/s2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
( time python bench.py ) > gpurun_out/s2/bench_full2.json 2> gpurun_out/s2/bench_full2.err
tail -3 gpurun_out/s2/bench_full2.err
python -c "
import json
d=json.load(open('gpurun_out/s2/bench_full2.json'))
print(d['value'], d['e2e'], d['roofline']['frac'], d['cpu_baseline'])
"
