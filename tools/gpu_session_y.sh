set -x
python -m pytest tests/test_gpu_tcgen05.py -q -x -k "layernorm_in" 2>&1 | tail -8 > gpurun_out/r2y_pytest.log
cat gpurun_out/r2y_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --other-configs none --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
grep "breakdown" gpurun_out/r2y_bench.err | head -40
FBANET_LN_IN_GEMM=0 timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --other-configs none --no-cpu-baseline > gpurun_out/r2y_bench_noln.json 2> gpurun_out/r2y_bench_noln.err
grep "breakdown" gpurun_out/r2y_bench_noln.err | head -12
python - <<'PY'
import json
for f in ("gpurun_out/r2y_bench.json","gpurun_out/r2y_bench_noln.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
PY
