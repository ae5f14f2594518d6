mkdir -p gpurun_out/r2d
python -m pytest tests -m gpu -q -x > gpurun_out/r2d/pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2d/pytest.log
for cl in 0 1 2 4; do echo "== CL $cl random"; SARPOST_NMS_CLUSTER=$cl SARPOST_LIB_PATH=$PWD/sar-yolo_b200/libsarpost_prof.so python tools/phase_prof.py cfg3 0; done > gpurun_out/r2d/phase_random.txt 2>&1
for cl in 0 1 8; do echo "== CL $cl blobs"; SARPOST_NMS_CLUSTER=$cl SARPOST_LIB_PATH=$PWD/sar-yolo_b200/libsarpost_prof.so python tools/phase_prof.py cfg3 50; done > gpurun_out/r2d/phase_blobs.txt 2>&1
cat gpurun_out/r2d/phase_random.txt
python tools/fuzz_parity.py 300 > gpurun_out/r2d/fuzz.txt 2>&1; tail -2 gpurun_out/r2d/fuzz.txt
