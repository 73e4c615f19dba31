set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r28; mkdir -p $O
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_w3.json 2> $O/bench_w3.err; tail -2 $O/bench_w3.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > $O/bench_w3_k20.json 2> $O/bench_w3_k20.err; tail -2 $O/bench_w3_k20.err
python - <<'PY'
import json
for f in ('bench_w3','bench_w3_k20'):
    d=json.load(open(f'gpurun_out/r28/{f}.json'))
    print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])
PY
