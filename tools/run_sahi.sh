N=${1:-2}
mkdir -p gpurun_out/r2s
for inf in 2 4; do
SARPOST_BENCH_SAHI_INFLIGHT=$inf python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$inf bench.py --gpus $N --workload cfg4 --steps 100 > gpurun_out/r2s/cfg4_x${N}_inf$inf.json 2> gpurun_out/r2s/cfg4_x${N}_inf$inf.err; echo "N=$N inflight=$inf rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2s/cfg4_x${N}_inf$inf.json').read().strip().splitlines()[-1])
    s=d['sahi']
    print('value %.0f tiles/s  ms %.4f  sharding %s' % (s['value'], s['ms_per_step'], s['sharding']))
    for k,v in s['by_sharding'].items():
        print(' ', k, 'launch', v['launch'], 'best %.4f ms' % v['ms_per_step'], 'eager %.4f' % v['eager']['ms_per_step'], 'graphs', v['graphs'], 'one %.4f' % v['one_in_flight']['ms_per_step'], {a:round(b,4) for a,b in v['phase_ms_max_over_ranks'].items()})
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2s/cfg4_x${N}_inf$inf.err').read()[-2500:])
PY
done
