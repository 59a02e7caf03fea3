#!/bin/bash
# round 2, GPU call 43: full GPU suite (geographic row strips, k_deps_dinf edge path), stencil timings, default bench line, launch list
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'passed\|failed\|Error\|error' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-1500))"; }
step tests_gpu 900 python -m pytest tests -m gpu -x -q
tail -5 gpurun_out/tests_gpu.log
step stencils_16384d 600 python scripts/stencil_bench.py 16384 7
grep -v "^{" gpurun_out/stencils_16384d.log | head -8
step bench_default 600 python bench.py
step launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench16384.csv python bench.py --size 16384 --steps 2 --warmup 1 --no-cpu --no-same-config
python - <<'PY'
import json
for l in open('gpurun_out/bench_default.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['value'], d['ms_per_step'], d['e2e']['value'], json.dumps(d['roofline'].get('per_kernel_ms')), json.dumps(d['roofline'].get('pipeline'))[:1500])
PY
