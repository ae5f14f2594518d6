N=${1:-2}
mkdir -p gpurun_out/r2mg
python -m pytest tests -m gpu -q -x -k "peer_memory or sharded" > gpurun_out/r2mg/pytest_mg_$N.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2mg/pytest_mg_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/peer_gather_check.py > gpurun_out/r2mg/peer_check_$N.log 2>&1; echo peer rc=$?; tail -4 gpurun_out/r2mg/peer_check_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/r2mg/bench_cfg3_x$N.json 2> gpurun_out/r2mg/bench_cfg3_x$N.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/r2mg/bench_cfg3_x$N.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'single', d['single_stream']['value'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='api'})
s=d['sahi']
print('sahi value %.0f tiles/s ms %.4f sharding %s' % (s['value'], s['ms_per_step'], s['sharding']))
for k,v in s['by_sharding'].items():
    print(' ', k, 'launch', v['launch'], 'best %.4f ms' % v['ms_per_step'], 'eager %.4f' % v['eager']['ms_per_step'], 'graphs', v['graphs'], 'one %.4f' % v['one_in_flight']['ms_per_step'], {a:round(b,4) for a,b in v['phase_ms_max_over_ranks'].items()})
print('clocks', d['clocks'])
PY
[ $N -le 2 ] && python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r2mg/bench_ref_x$N.json 2> gpurun_out/r2mg/bench_ref_x$N.err; echo ref rc=$?; tail -c 600 gpurun_out/r2mg/bench_ref_x$N.json
