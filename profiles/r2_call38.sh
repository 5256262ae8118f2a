mkdir -p gpurun_out
timeout 120 python profiles/gemm_latency.py 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench:', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['us_per_launch'])"
