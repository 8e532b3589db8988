#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_schnet.py tests/test_gpu_config_parity.py tests/test_gpu_tc.py -q -k "schnet or config2 or graphed" 2>&1 | tail -2
python bench.py --only 2 --steps 20 --warmup 5 --no-cpu-baseline --no-strict 2>/dev/null | python -c "
import json,sys
s=sys.stdin.read(); d=json.loads(s[s.find('{\"metric'):].splitlines()[0]); print('ms', round(d['ms_per_step'],4), 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['graph_kernel_nodes'])"
