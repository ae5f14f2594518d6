mkdir -p gpurun_out/r2n
python -m pytest tests -m gpu -q -x > gpurun_out/r2n/pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2n/pytest.log
python tools/fuzz_parity.py 500 51 2>&1 | tail -1
for cl in 1 2 8; do SARPOST_NMS_CLUSTER=$cl python tools/fuzz_parity.py 150 6$cl 2>&1 | tail -1; done
python sar-yolo_b200/build.py --prof > /dev/null
for b in 0 50; do SARPOST_LIB_PATH=$PWD/sar-yolo_b200/libsarpost_prof.so python tools/phase_prof.py cfg3 $b; done > gpurun_out/r2n/phase.txt 2>&1; cat gpurun_out/r2n/phase.txt
SARPOST_LIB_PATH=$PWD/sar-yolo_b200/libsarpost_prof.so python tools/phase_prof.py cfg1 0 | tail -16
for w in cfg3 cfg1; do python bench.py --workload $w --quick --steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$w value %.0f ms %.4f single %.4f' % (d['value'], d['ms_per_step'], d['single_stream']['ms_per_step']), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})
"; done
