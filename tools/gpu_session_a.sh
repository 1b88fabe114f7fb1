set -x
nvidia-smi -L
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/run_train_step.py --dtype bf16 --steps 3 --warmup 2 > gpurun_out/r2a_train_bf16_2gpu.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/run_train_step.py --dtype fp32 --steps 3 --warmup 2 > gpurun_out/r2a_train_fp32_2gpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2a_bench_2gpu.json 2> gpurun_out/r2a_bench_2gpu.err
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown > gpurun_out/r2a_bench_1gpu.json 2> gpurun_out/r2a_bench_1gpu.err
tail -3 gpurun_out/r2a_pytest.log
