N=${1:-8}
mkdir -p gpurun_out/r2s
python -m pytest tests -m gpu -q -x -k "peer_memory or sharded" 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/peer_gather_check.py 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --workload cfg4 --steps 100 > gpurun_out/r2s/cfg4_x${N}.json 2> gpurun_out/r2s/cfg4_x${N}.err; echo "N=$N rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2s/cfg4_x${N}.json').read().strip().splitlines()[-1])
s=d['sahi']
print('value %.0f tiles/s  ms %.4f  sharding %s' % (s['value'], s['ms_per_step'], s['sharding']))
for k,v in s['by_sharding'].items():
    print(' ', k, 'launch', v['launch'], 'best %.4f ms' % v['ms_per_step'], 'eager %.4f' % v['eager']['ms_per_step'], 'graphs', v['graphs'], 'one %.4f' % v['one_in_flight']['ms_per_step'], {a:round(b,4) for a,b in v['phase_ms_max_over_ranks'].items()})
PY
