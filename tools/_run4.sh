python -m pytest tests/test_rx_chain_gpu.py tests/test_ldpc_gpu.py tests/test_rtl_fsk_tool.py tests/test_cli_pipeline.py -x -q -m gpu > gpurun_out/pt4.log 2>&1
tail -15 gpurun_out/pt4.log
python tools/bench_secondary.py > gpurun_out/secondary.json 2> gpurun_out/secondary.err
tail -5 gpurun_out/secondary.err
head -c 3000 gpurun_out/secondary.json
