mkdir -p gpurun_out/r2a
python -m pytest tests -m gpu -q -x > gpurun_out/r2a/pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2a/pytest.log
python bench.py > gpurun_out/r2a/bench_cfg3.json 2> gpurun_out/r2a/bench_cfg3.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a/bench_ref.json 2> gpurun_out/r2a/bench_ref.err; echo ref rc=$?
for w in cfg1 cfg2 cfg5; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2a/bench_$w.json 2> gpurun_out/r2a/bench_$w.err; echo $w rc=$?; done
python tools/worstcase_probe.py > gpurun_out/r2a/worstcase.txt 2>&1; cat gpurun_out/r2a/worstcase.txt
python -c "
import json
for w in ['cfg3','cfg1','cfg2','cfg5']:
    d=json.loads(open('gpurun_out/r2a/bench_%s.json'%w).read().strip().splitlines()[-1])
    print(w, 'value %.0f ms %.4f single %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['roofline']['frac']), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()}, 'e2e', d.get('e2e',{}).get('value'))
"
