set -x
timeout 300 python -m pytest tests/test_gpu_tcgen05.py -q -k "faf_fuse" 2>&1 | tail -30 > gpurun_out/r2j_pytest_faf.log
tail -12 gpurun_out/r2j_pytest_faf.log
if grep -q "failed\|rror" gpurun_out/r2j_pytest_faf.log; then echo FAF_FAILED; else
timeout 600 python -m pytest tests/test_gpu_model.py -q 2>&1 | tail -6 > gpurun_out/r2j_pytest_model.log
tail -3 gpurun_out/r2j_pytest_model.log
timeout 300 python bench.py --steps 10 --warmup 3 --breakdown --other-configs none --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
grep "faf\|896\|64->16 @160" gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_narrow_io']['value'])
PY
fi
