set -x
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_final_bench_8gpu.json 2> gpurun_out/r2_final_bench_8gpu.err
tail -3 gpurun_out/r2_final_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_8gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_blocking']['value'], d['e2e_narrow_io']['value'], d['clocks'])
o=d['other_configs']; print({k:(v.get('ms_per_step') or v.get('ms_per_image')) for k,v in o.items()})
PY
