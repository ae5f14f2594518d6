mkdir -p gpurun_out/r2h
for ldg in 0 1; do echo "== SARPOST_K1_FORCE_LDG=$ldg"
for w in cfg1 cfg5 cfg2 cfg3; do SARPOST_K1_FORCE_LDG=$ldg python bench.py --workload $w --quick --steps 300 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$w value %.0f ms %.4f single %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['roofline']['frac']), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})
"; done; done 2>&1 | tee gpurun_out/r2h/k1_ldg_vs_tma.txt
