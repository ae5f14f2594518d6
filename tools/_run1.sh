mkdir -p gpurun_out/r2i
python bench.py > gpurun_out/r2i/bench.json 2> gpurun_out/r2i/bench.err; echo bench rc=$?; tail -3 gpurun_out/r2i/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i/bench.json'))
print('value', d['value'], d['ms_per_step'], 'launches', d['launches_per_step']); print('two', d['two_streams']); print('single', d['single_stream'])
print(d['roofline']['frac'], d['roofline']['stage_ms'])
print('cat', json.dumps(d['cat_layout'])[:900])
print('clustered', {k:d['clustered'][k] for k in ('value','ms_per_step','single_stream','k1_ms','k4_ms','k5_ms')})
print('refgpu', d['reference_gpu']['value'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_h2d_ceiling'], 'cpu', d['cpu_baseline']['value'])
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2i/ref_arm.json 2> gpurun_out/r2i/ref_arm.err; tail -c 600 gpurun_out/r2i/ref_arm.json
