for a in "c1 1251 0.25" "c1 2600 0.25" "c4 1000 0.25" "c4 2600 0.25"; do timeout 60 python tools/dbg_multi.py $a 2>&1 | tail -4 | cut -c1-250; echo "rc=$?"; done
PIRIP_FSK_NO_CHAIN=1 timeout 60 python tools/dbg_multi.py c4 2600 0.25 2>&1 | tail -4 | cut -c1-250
timeout 300 python -m pytest tests/test_ldpc_gpu.py tests/test_rx_chain_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/bench_extra.py --only ldpc > gpurun_out/ldpc_r02c.json 2> gpurun_out/ldpc_c.err; python - <<'PY'
import json
for l in open('gpurun_out/ldpc_r02c.json'):
    d=json.loads(l); print(d['mode'], round(d['value']/1e6,2),'Mcw/s', d['ms_per_step'], d['mean_iterations'], d.get('parity_vs_oracle(bits,iter,pcc)'), d.get('payload_equal_to_exact_mode_on_converged_words'))
PY
