run() { echo "== $*"; timeout 60 "$@" 2>&1 | tail -3 | cut -c1-200; }
run python tools/dbg_multi.py c1 1400 0.25
run python tools/dbg_multi.py c1 2600 0.25
run python tools/dbg_multi.py c4 2600 0.25
timeout 300 python -m pytest tests/test_ldpc_gpu.py tests/test_rx_chain_gpu.py tests/test_fsk_gpu_api.py tests/test_rtl_fsk_tool.py -x -q -m gpu --timeout 200 2>&1 | tail -8
timeout 120 python tools/bench_extra.py --only ldpc > gpurun_out/ldpc_r02d.json 2> gpurun_out/ldpc_d.err; python - <<'PY'
import json
for l in open('gpurun_out/ldpc_r02d.json'):
    d=json.loads(l); print(d['mode'], round(d['value']/1e6,2),'Mcw/s', d['ms_per_step'], d['mean_iterations'], d.get('parity_vs_oracle(bits,iter,pcc)'), d.get('payload_equal_to_exact_mode_on_converged_words'))
PY
timeout 300 python tools/bench_secondary.py --only chain_c4 > gpurun_out/sec_c4.json 2> gpurun_out/sec_c4.err; echo rc=$?; tail -3 gpurun_out/sec_c4.err
timeout 300 python tools/bench_secondary.py --only c3_4fsk_llr > gpurun_out/sec_c3.json 2> gpurun_out/sec_c3.err; echo rc=$?; tail -2 gpurun_out/sec_c3.err
