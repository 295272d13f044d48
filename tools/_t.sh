python tools/prof_fsk.py --captures 1250 --seconds 1.0 --iters 5 2>&1 | tail -2 | cut -c1-120
timeout 300 python -m pytest tests/test_fsk_fast_gpu.py -x -q -m gpu 2>&1 | tail -2
