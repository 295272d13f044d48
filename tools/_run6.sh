timeout 300 python tools/bench_secondary.py --only chain_c4 > gpurun_out/sec_c4.json 2> gpurun_out/sec_c4.err; echo rc=$?; tail -8 gpurun_out/sec_c4.err
timeout 300 python tools/bench_secondary.py --only c3_4fsk_llr > gpurun_out/sec_c3.json 2> gpurun_out/sec_c3.err; echo rc=$?; tail -4 gpurun_out/sec_c3.err
timeout 600 python -m pytest tests/test_fsk_gpu_api.py tests/test_rtl_fsk_tool.py tests/test_cli_pipeline.py tests/test_fsk_gpu_batch_kernel.py -x -q -m gpu --timeout 120 2>&1 | tail -15
