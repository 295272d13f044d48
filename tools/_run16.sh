timeout 100 python tools/dbg_c4.py 5000 2>&1 | tail -1
PIRIP_FSK_FAST_NO_RERUN=1 timeout 100 python tools/dbg_c4.py 5000 2>&1 | tail -1
PIRIP_FSK_NO_CHAIN=1 PIRIP_FSK_FAST_NO_RERUN=1 PIRIP_FSK_FAST_NO_EXIT=1 timeout 100 python tools/dbg_c4.py 5000 2>&1 | tail -1
timeout 100 python tools/dbg_c4.py 1251 2>&1 | tail -1
