python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err; echo rc=$?; tail -3 gpurun_out/bench_r02_2gpu.err; cut -c1-300 gpurun_out/bench_r02_2gpu.json
timeout 300 python -m pytest tests/test_fsk_gpu_api.py -x -q -m gpu -k multi 2>&1 | tail -3
