set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_final_pytest.log; cat gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -2 gpurun_out/r2_final_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err; cut -c1-300 gpurun_out/r2_final_bench_ref.json
timeout 900 python bench.py --steps 10 --warmup 3 --breakdown > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
grep "breakdown" gpurun_out/r2_final_bench.err | head -8
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_blocking']['value'], d['e2e_narrow_io']['value'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])
print([ (k['kernel'][:28], round(k['frac'],3)) for k in d['hbm_kernels']])
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:fbanet -c 1300 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --other-configs none --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1
grep -c head_conv gpurun_out/r2f_launches.csv
