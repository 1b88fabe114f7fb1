set -x
nvidia-smi -L
python -m pytest tests/test_gpu_model.py -q -k "two_gpus" 2>&1 | tail -3 > gpurun_out/r2_final_pytest_2gpu.log; cat gpurun_out/r2_final_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_final_bench_2gpu.json 2> gpurun_out/r2_final_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_blocking']['value'], d['e2e_narrow_io']['value'], d['clocks'])
print(json.dumps(d['other_configs'])[:1500])
PY
