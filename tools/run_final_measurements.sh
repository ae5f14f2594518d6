O=gpurun_out/r2final; mkdir -p $O
python -m pytest tests -m gpu -q -x -rs > $O/pytest.log 2>&1; echo pytest rc=$?; tail -4 $O/pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > $O/reference_arm_cpu.json 2> $O/reference_arm_cpu.err; echo ref rc=$?
python bench.py > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo bench rc=$?
python bench.py --steps 20 --warmup 5 > $O/bench_cfg3_driver_args.json 2> $O/bench_cfg3_driver_args.err; echo bench20 rc=$?
for w in cfg1 cfg2 cfg5; do python bench.py --workload $w --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo $w rc=$?; done
python tools/worstcase_probe.py > $O/worstcase_final.txt 2>&1; cat $O/worstcase_final.txt
python tools/host_overhead_probe.py cfg1 cfg5 cfg3 > $O/host_overhead_per_call.txt 2>&1
python sar-yolo_b200/build.py --prof > /dev/null
for a in "cfg3 0" "cfg3 50" "cfg1 0"; do set -- $a; SARPOST_LIB_PATH=$PWD/sar-yolo_b200/libsarpost_prof.so python tools/phase_prof.py $1 $2 > $O/k4_phase_cycles_$1_blobs$2.txt 2>&1; done
python - <<PY
import json
for w in ['cfg3','cfg3_driver_args','cfg1','cfg2','cfg5']:
    d=json.loads(open('$O/bench_%s.json'%w).read().strip().splitlines()[-1])
    print(w, 'value %.0f ms %.4f (%s) single %.4f general %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['config']['value_is'], d['single_stream']['ms_per_step'], d['single_stream_general']['ms_per_step'], d['roofline']['frac']), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()}, 'e2e', round(d.get('e2e',{}).get('value') or 0), 'clustered', round((d.get('clustered') or {}).get('value') or 0), 'refgpu', round((d.get('reference_gpu') or {}).get('value') or 0,1), 'sahi', round((d.get('sahi') or {}).get('value') or 0))
PY
