set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2r_pytest.log
tail -3 gpurun_out/r2r_pytest.log
python tools/prof_ops.py --case attn_enc0_64x1_s5,attn_enc1_128x2_s5,attn_enc0_64x1_s0 --reps 5 > gpurun_out/r2r_prof.log 2>&1
cat gpurun_out/r2r_prof.log
timeout 300 python bench.py --steps 10 --warmup 3 --breakdown --other-configs none --no-cpu-baseline > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
grep "attention\|faf\|64->16" gpurun_out/r2r_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2r_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_narrow_io']['value'])
PY
