mkdir -p gpurun_out/r2j
python -m pytest tests -m gpu -q -x -rs > gpurun_out/r2j/pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2j/pytest.log
python tools/fuzz_parity.py 600 76 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
