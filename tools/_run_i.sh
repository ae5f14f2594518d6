mkdir -p gpurun_out/r2i
python -m pytest tests -m gpu -q -x > gpurun_out/r2i/pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2i/pytest.log
python tools/fuzz_parity.py 400 75 2>&1 | tail -1
for w in cfg1 cfg2 cfg3; do python bench.py --workload $w --quick --steps 300 2>gpurun_out/r2i/q_$w.err > gpurun_out/r2i/q_$w.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2i/q_$w.json').read())
print('$w value %.0f ms %.4f single(plan) %.4f general %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['single_stream_general']['ms_per_step'], d['roofline']['frac']), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()}, d['config']['value_is'])
" || tail -5 gpurun_out/r2i/q_$w.err; done
python bench.py --workload cfg3 --no-split --quick --steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('cfg3 cat-layout value %.0f single %.4f' % (d['value'], d['single_stream']['ms_per_step']), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})"
