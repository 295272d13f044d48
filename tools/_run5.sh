timeout 300 python -m pytest tests/test_ldpc_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/bench_extra.py --only ldpc > gpurun_out/ldpc_r02b.json 2> gpurun_out/ldpc_b.err; python - <<'PY'
import json
for l in open('gpurun_out/ldpc_r02b.json'):
    d=json.loads(l); print(d['mode'], round(d['value']/1e6,2),'Mcw/s', d['ms_per_step'], d['mean_iterations'], d.get('parity_vs_oracle(bits,iter,pcc)'), d.get('payload_equal_to_exact_mode_on_converged_words'))
PY
timeout 200 python tools/bench_secondary.py --only chain_c4 --ncap 500 > gpurun_out/sec_c4.json 2> gpurun_out/sec_c4.err; tail -12 gpurun_out/sec_c4.err
timeout 200 python tools/bench_secondary.py --only c3_4fsk_llr --ncap 150 > gpurun_out/sec_c3.json 2> gpurun_out/sec_c3.err; tail -4 gpurun_out/sec_c3.err
timeout 400 python tools/bench_secondary.py --only c1_single_stream > gpurun_out/sec_c1.json 2> gpurun_out/sec_c1.err; tail -4 gpurun_out/sec_c1.err
