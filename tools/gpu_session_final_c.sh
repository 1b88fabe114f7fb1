set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_final_pytest.log; cat gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -2 gpurun_out/r2_final_smoke.log
python tools/prof_ops.py --case mlp16_dec1_128,mlp16_enc1_128,mlp16_enc0_64,mlp_dec1_128,mlp_enc1_128,mlp_enc0_64 --reps 10 > gpurun_out/r2_final_mlp_prof.log 2>&1; cat gpurun_out/r2_final_mlp_prof.log
timeout 900 python bench.py --steps 10 --warmup 3 --breakdown > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
grep "breakdown" gpurun_out/r2_final_bench.err | head -16
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_narrow_io']['value'], d['roofline']['frac'], d['clocks'], d['cpu_baseline'])
PY
