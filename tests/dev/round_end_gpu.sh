# one GPU call: full GPU test-suite, smoke, the default bench (+ one variant)
set -x; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 2500 gpurun_out/bench_default.json
ZKFL_WITNESS_COOP=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-msm 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('coop forced:', round(d['value'],1), d['stages_ms']['witness'])"
