PIRIP_FSK_CHAIN_DIAG=1 python tools/prof_fsk.py --captures 1250 --seconds 1.0 --iters 5 2>&1 | grep -E "captures|side chain" | cut -c1-200
timeout 500 python -m pytest tests/test_fsk_fast_gpu.py tests/test_rx_chain_gpu.py -x -q -m gpu 2>&1 | tail -2
