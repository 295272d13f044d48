timeout 300 python tools/dbg_seeds.py 3 4 7 2>&1 | grep seed | cut -c1-200
timeout 600 python -m pytest tests/test_fsk_fast_gpu.py tests/test_lib_abi.py -x -q --timeout 300 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_r02_4gpu_d.json 2> gpurun_out/bench_r02_4gpu_d.err; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_4gpu_d.json')); print(d['value'], d['ms_per_step'], d['config'].get('gather')[:40], d['config'].get('gather_equals_nccl_all_gather'), d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_ceiling']['ms_per_step'], d['e2e'].get('host_affinity'), d['fast_mode'])"
