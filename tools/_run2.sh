python -m pytest tests/test_fsk_fast_gpu.py tests/test_rx_chain_gpu.py -x -q -m gpu > gpurun_out/pt_fast.log 2>&1
tail -5 gpurun_out/pt_fast.log
python tools/prof_fsk.py --captures 1250 --seconds 1.0 --iters 4 2>&1 | tail -2
PIRIP_FSK_FAST_NO_SPEC=1 python tools/prof_fsk.py --captures 1250 --seconds 1.0 --iters 4 2>&1 | tail -2
PIRIP_FSK_FAST_NO_MERGE=1 python tools/prof_fsk.py --captures 1250 --seconds 1.0 --iters 4 2>&1 | tail -2
ncu --set full --import-source on -k regex:fsk_fast --launch-skip 2 -c 1 -o gpurun_out/fast_v4 -f python tools/prof_fsk.py --captures 1250 --seconds 0.25 --iters 3 > gpurun_out/ncu_fast4.log 2>&1
tail -3 gpurun_out/ncu_fast4.log
