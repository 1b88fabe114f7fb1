set -x
nvidia-smi -L
python -m pytest tests/test_gpu_model.py -q -k "two_gpus" 2>&1 | tail -3 > gpurun_out/r2g_pytest_2gpu.log; cat gpurun_out/r2g_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err
tail -c 3000 gpurun_out/r2g_bench_2gpu.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err; cat gpurun_out/r2g_bench_ref.json | cut -c1-400
