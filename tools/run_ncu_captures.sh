mkdir -p gpurun_out/r2ncu
python bench.py --quick --streams 1 --steps 3 --warmup 3 > gpurun_out/r2ncu/plain.json 2>/dev/null; echo plain rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ncu/launches_cfg3.csv python bench.py --quick --streams 1 --steps 3 --warmup 3 > gpurun_out/r2ncu/ncu_launch.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k1_fused|k4_nms|k5_gather" -s 9 -c 3 -o gpurun_out/r2ncu/full_cfg3 python bench.py --quick --streams 1 --steps 3 --warmup 3 > gpurun_out/r2ncu/ncu_full.log 2>&1; echo full rc=$?
python tools/fp16_probe.py half 20 > /dev/null 2>&1; echo fp16 rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k1_fused" -s 10 -c 1 -o gpurun_out/r2ncu/k1_fp16_cfg3 python tools/fp16_probe.py half 5 > gpurun_out/r2ncu/ncu_fp16.log 2>&1; echo fp16 ncu rc=$?
