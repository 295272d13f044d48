run() { echo "== $*"; timeout 40 env "$@" 2>&1 | tail -3 | cut -c1-200; }
run A=1 python tools/dbg_multi.py c1 1400 0.25
run PIRIP_FSK_FAST_NO_RERUN=1 python tools/dbg_multi.py c1 1400 0.25
run PIRIP_FSK_FAST_NO_RERUN=1 PIRIP_FSK_NO_CHAIN=1 python tools/dbg_multi.py c1 1400 0.25
run PIRIP_FSK_FAST_NO_RERUN=1 PIRIP_FSK_NO_CHAIN=1 PIRIP_FSK_FAST_NO_EXIT=1 python tools/dbg_multi.py c1 1400 0.25
run PIRIP_FSK_FAST_NO_RERUN=1 PIRIP_FSK_NO_TMA=1 python tools/dbg_multi.py c1 1400 0.25
run PIRIP_FSK_MODE=exact python tools/dbg_multi.py c1 1400 0.25
