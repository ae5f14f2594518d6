mkdir -p gpurun_out/r2b
python -m pytest tests -m gpu -q -x > gpurun_out/r2b/pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2b/pytest.log
python tools/fuzz_parity.py 300 71 2>&1 | tail -1
for w in cfg5 cfg2 cfg3; do python bench.py --workload $w --quick --steps 200 2>/dev/null > gpurun_out/r2b/q_$w.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2b/q_$w.json').read())
print('$w value %.0f ms %.4f single %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['roofline']['frac']), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})
"; done
python bench.py --workload cfg5 --batch 256 --quick --steps 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('cfg5 b256 value %.0f ms %.4f single %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['roofline']['frac']), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})
"
ncu --set full --clock-control none --import-source on -k regex:"k1_fused" -c 2 -o gpurun_out/r2b/k1_cfg5 python bench.py --workload cfg5 --quick --streams 1 --steps 3 --warmup 3 > gpurun_out/r2b/ncu_cfg5.log 2>&1; echo ncu rc=$?
