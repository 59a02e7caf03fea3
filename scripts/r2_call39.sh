#!/bin/bash
# round 2, GPU call 39: the new sibling tools on the GPU, k_fill_init on the tile ring (timings), full GPU suite
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'passed\|failed\|smoke\|Error\|error' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-2500))"; }
step tests_new 600 python -m pytest tests/test_gpu_parity.py -x -q -k "conc_lim or slopearea or golden or depression or live_reference or row_strip"
step tests_gpu 900 python -m pytest tests -m gpu -x -q
step bench_16384 600 python bench.py --size 16384 --no-cpu
python - <<'PY'
import json
for l in open('gpurun_out/bench_16384.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print(json.dumps(d.get('roofline', {}).get('pipeline', d.get('roofline')))[:3000])
        print(d.get('value'), d.get('ms_per_step'))
PY
