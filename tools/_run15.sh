timeout 600 python -m pytest tests/test_fsk_fast_gpu.py tests/test_rx_chain_gpu.py tests/test_fsk_gpu.py -x -q -m gpu --timeout 200 2>&1 | tail -4
timeout 300 python tools/bench_secondary.py --only chain_c4 > gpurun_out/sec_c4.json 2> gpurun_out/sec_c4.err; echo rc=$?
timeout 300 python tools/bench_secondary.py --only c3_4fsk_llr > gpurun_out/sec_c3.json 2> gpurun_out/sec_c3.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/sec_c4.json'))['chain_c4']
print('c4', d['value'], d['ms_per_step'], {k:(v['value'],v['ms_demodulator_alone']) for k,v in d['modes'].items()})
d=json.load(open('gpurun_out/sec_c3.json'))['c3_4fsk_llr']
print('c3', d['value'], d['ms_per_step'], d['parity'])
PY
timeout 200 python tools/bench_extra.py --only chain 2>/dev/null | cut -c1-600
