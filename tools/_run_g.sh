mkdir -p gpurun_out/r2g
python tools/host_overhead_probe.py cfg1 cfg5 cfg3 2>&1 | tee gpurun_out/r2g/host_overhead.txt
for cs in "3 2" "4 1" "5 1" "6 1"; do set -- $cs; echo "== K1 (TMA) ctas=$1 stages=$2"
for w in cfg5 cfg2 cfg3; do SARPOST_K1_CTAS=$1 SARPOST_K1_STAGES=$2 python bench.py --workload $w --quick --steps 300 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$w value %.0f ms %.4f single %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['roofline']['frac']), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})
"; done; done 2>&1 | tee gpurun_out/r2g/k1_sweep_tma.txt
python tools/fp16_probe.py 2>&1 | tail -12 | tee gpurun_out/r2g/fp16.txt
