set -x
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --breakdown --other-configs none --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
FBANET_FUSE_MLP=0 timeout 300 python bench.py --steps 10 --warmup 3 --other-configs none --no-cpu-baseline > gpurun_out/r2h_bench_nofuse.json 2> gpurun_out/r2h_bench_nofuse.err
python - <<'PY'
import json
for f in ('r2h_bench','r2h_bench_nofuse'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
PY
