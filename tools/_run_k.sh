mkdir -p gpurun_out/r2k
python -m pytest tests -m gpu -q -x > gpurun_out/r2k/pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2k/pytest.log
python tools/fuzz_parity.py 400 77 2>&1 | tail -1
for res in 0 1; do echo "== SARPOST_PIPE_NO_RESERVE=$res"
for w in cfg3 cfg2 cfg5; do SARPOST_PIPE_NO_RESERVE=$res python bench.py --workload $w --quick --steps 400 2>gpurun_out/r2k/q_$w.err > gpurun_out/r2k/q_${w}_$res.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2k/q_${w}_$res.json').read())
print('$w value %.0f ms %.4f single(plan) %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['roofline']['frac']), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()}, d['config']['value_is'], 'pipe rows same', d['config']['in_flight'][-60:])
" || tail -5 gpurun_out/r2k/q_$w.err; done; done
for m in 2; do echo "== RESERVE_MULT=$m"; for w in cfg3 cfg5; do SARPOST_PIPE_RESERVE_MULT=$m python bench.py --workload $w --quick --steps 400 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$w value %.0f ms %.4f' % (d['value'], d['ms_per_step']))"; done; done
