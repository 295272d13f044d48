set -x
python -m pytest tests/test_rx_chain_gpu.py tests/test_ldpc_gpu.py -x -q -m gpu > gpurun_out/pt_rx.log 2>&1
tail -5 gpurun_out/pt_rx.log
python -m pytest tests -x -q -m gpu --deselect tests/test_rx_chain_gpu.py --deselect tests/test_ldpc_gpu.py > gpurun_out/pt_all2.log 2>&1
tail -3 gpurun_out/pt_all2.log
ncu --set full --import-source on -k regex:fsk_fast --launch-skip 2 -c 1 -o gpurun_out/fast_v3 -f python tools/prof_fsk.py --captures 1250 --seconds 0.25 --iters 3 > gpurun_out/ncu_fast3.log 2>&1
tail -3 gpurun_out/ncu_fast3.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err
cat gpurun_out/bench_r02b.json | cut -c1-600
