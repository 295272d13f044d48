timeout 300 python -m pytest tests/test_ldpc_gpu.py tests/test_rx_chain_gpu.py -x -q -m gpu --timeout 200 2>&1 | tail -4
timeout 120 python tools/bench_extra.py --only ldpc > gpurun_out/ldpc_r02e.json 2> gpurun_out/ldpc_e.err; python - <<'PY'
import json
for l in open('gpurun_out/ldpc_r02e.json'):
    d=json.loads(l); print(d['mode'], round(d['value']/1e6,2),'Mcw/s', d['ms_per_step'], d['mean_iterations'], d.get('parity_vs_oracle(bits,iter,pcc)'), d.get('payload_equal_to_exact_mode_on_converged_words'), d['word_error_rate'])
PY
ncu --set full --import-source on -k regex:ldpc_decode_warp --launch-skip 20 -c 1 -o gpurun_out/ldpc_ms -f python tools/bench_extra.py --only ldpc > gpurun_out/ncu_ldpc.log 2>&1; tail -2 gpurun_out/ncu_ldpc.log
