timeout 300 python tools/dbg_seeds.py 0 1 2 3 4 5 6 7 2>&1 | grep seed
