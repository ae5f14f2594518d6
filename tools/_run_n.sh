for dpt in 2 3 4; do echo "== depth $dpt"
for args in "--workload cfg3" "--workload cfg3 --blobs 50" "--workload cfg5" "--workload cfg2"; do SARPOST_BENCH_PIPE_DEPTH=$dpt python bench.py $args --quick --steps 300 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$args value %.0f ms %.4f (%s) single %.4f general %.4f' % (d['value'], d['ms_per_step'], d['config']['value_is'], d['single_stream']['ms_per_step'], d['single_stream_general']['ms_per_step']))"; done; done
