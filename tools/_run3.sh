python -m pytest tests/test_fsk_fast_gpu.py -x -q -m gpu > gpurun_out/pt_fast.log 2>&1
tail -3 gpurun_out/pt_fast.log
python tools/prof_fsk.py --captures 1250 --seconds 1.0 --iters 4 2>&1 | tail -2 | cut -c1-200
PIRIP_FSK_FAST_RO_EVEN=1 python tools/prof_fsk.py --captures 1250 --seconds 1.0 --iters 4 2>&1 | tail -2 | cut -c1-100
