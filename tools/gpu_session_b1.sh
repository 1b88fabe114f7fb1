set -x
timeout 900 python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/r2b1_bench.json 2> gpurun_out/r2b1_bench.err
grep "breakdown" gpurun_out/r2b1_bench.err | head -24
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b1_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_narrow_io']['value'], d['roofline']['frac']); print(json.dumps(d.get('other_configs'))[:600])
PY
