timeout 600 python -m pytest tests/test_ldpc_gpu.py tests/test_rx_chain_gpu.py tests/test_ldpc_cli.py tests/test_fsk_gpu.py -x -q -m gpu --timeout 200 2>&1 | tail -5
timeout 120 python tools/bench_extra.py --only ldpc > gpurun_out/ldpc_r02g.json 2> gpurun_out/ldpc_g.err; python - <<'PY'
import json
for l in open('gpurun_out/ldpc_r02g.json'):
    d=json.loads(l); print(d['mode'], round(d['value']/1e6,2),'Mcw/s', d['ms_per_step'], d['mean_iterations'], d.get('parity_vs_oracle(bits,iter,pcc)'), d.get('payload_equal_to_exact_mode_on_converged_words'), d['word_error_rate'])
PY
