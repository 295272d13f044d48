timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo rc=$?; tail -5 gpurun_out/bench_r02c.err; cut -c1-400 gpurun_out/bench_r02c.json
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pt_all3.log 2>&1; tail -5 gpurun_out/pt_all3.log
