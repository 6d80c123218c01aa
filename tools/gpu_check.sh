#!/bin/bash
# tools/gpu_check.sh [pytest targets...] -- run ON THE GPU BOX (under gpurun): the given GPU tests, then a short
# front-end bench with the per-kernel shares printed.
set -u
TARGETS=${@:-tests}
python -m pytest $TARGETS -m gpu -x -q 2>&1 | tail -5
python bench.py --no-ba --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
print(json.dumps(d.get('kernels', d.get('kernel_share', {})))[:2500])
print(json.dumps(d.get('roofline')))"
