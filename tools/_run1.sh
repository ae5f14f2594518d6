mkdir -p gpurun_out/r2g
for cfg in "3 2" "2 2" "2 3" "1 4" "1 6" "4 1"; do set -- $cfg; echo "== K1 ctas=$1 stages=$2"; SARPOST_K1_CTAS=$1 SARPOST_K1_STAGES=$2 python bench.py --quick --steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f single %.0f  ms %.4f single_ms %.4f stage %s' % (d['value'], d['single_stream']['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()}))
"; done 2>&1 | tee gpurun_out/r2g/k1_sweep.txt
