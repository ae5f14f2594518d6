mkdir -p gpurun_out/r2d
python -m pytest tests -m gpu -q -x > gpurun_out/r2d/pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r2d/pytest.log
python tools/fuzz_parity.py 300 72 2>&1 | tail -1
for w in cfg1 cfg5 cfg2 cfg3; do python bench.py --workload $w --quick --steps 300 2>gpurun_out/r2d/q_$w.err > gpurun_out/r2d/q_$w.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2d/q_$w.json').read())
print('$w value %.0f ms %.4f single(plan) %.4f general %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['single_stream_general']['ms_per_step'], d['roofline']['frac']), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()}, d['single_stream'].get('rows_identical_to_general_call'))
" || tail -5 gpurun_out/r2d/q_$w.err; done
python sar-yolo_b200/build.py --prof > /dev/null
for w in cfg1 cfg3; do SARPOST_LIB_PATH=$PWD/sar-yolo_b200/libsarpost_prof.so python tools/phase_prof.py $w 0; done > gpurun_out/r2d/phase.txt 2>&1; cat gpurun_out/r2d/phase.txt
