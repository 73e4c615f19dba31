set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r27; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -k "graph" 2>&1 | tail -5 > $O/pytest_graph.log; cat $O/pytest_graph.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -5 $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r27/bench.json'))
print(d['value'], d['e2e']['value'], d['gpu_launches'], {k:v.get('gpu_launches') for k,v in d['workloads'].items()})
PY
