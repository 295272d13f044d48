# Round-2 evidence sweep (one GPU): everything profiles/r02_* is made from.  Run under gpurun from the repo root.
set -x
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; cut -c1-200 gpurun_out/r02_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fsk|fir_dec|ldpc|pack_bits" -c 200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsk_fast --launch-skip 2 -c 1 -o gpurun_out/r02_fast -f python tools/prof_fsk.py --captures 1250 --seconds 1.0 --iters 3 > gpurun_out/r02_fast_ncu.log 2>&1; tail -2 gpurun_out/r02_fast_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode_warp --launch-skip 20 -c 1 -o gpurun_out/r02_ldpc -f python tools/bench_extra.py --only ldpc > gpurun_out/r02_ldpc_ncu.log 2>&1; tail -2 gpurun_out/r02_ldpc_ncu.log
timeout 600 python tools/bench_extra.py > gpurun_out/r02_extra_kernels.json 2> gpurun_out/r02_extra_kernels.err; cut -c1-300 gpurun_out/r02_extra_kernels.json
