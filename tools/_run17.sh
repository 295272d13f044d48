timeout 600 python -m pytest tests/test_fsk_fast_gpu.py tests/test_rx_chain_gpu.py -x -q -m gpu --timeout 300 2>&1 | tail -4
timeout 100 python tools/dbg_c4.py 5000 2>&1 | tail -1
timeout 200 python -c "
import __graft_entry__ as g
g.smoke()" 2>&1 | tail -2
