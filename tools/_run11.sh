for v in 4 8 16 32; do
LDPC_B200_PIPE_DIV=$v python bench.py --no-extra --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('div=$v', 'e2e', round(d['e2e']['value'],3), d['e2e']['matches_device_path'])"
done
