timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
nproc; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA node\(s\)" 
for v in 1 0; do
LDPC_B200_HOST_NARROW=$v python bench.py --no-extra --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('narrow=$v', 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],3), d['e2e']['matches_device_path'])"
done
for t in 4 8 12; do
LDPC_B200_HOST_THREADS=$t python bench.py --no-extra --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('threads=$t', 'e2e', round(d['e2e']['value'],3))"
done
