python -m pytest tests/test_rqs_gpu.py tests/test_chain_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r3d_tests.log
python scripts/rqs_bwd_ab.py 10:40 10:35 > gpurun_out/r3d_bwd_ab.log 2>&1
python scripts/rqs64_ab.py > gpurun_out/r3d_fwd_ab.log 2>&1
cat gpurun_out/r3d_tests.log gpurun_out/r3d_bwd_ab.log gpurun_out/r3d_fwd_ab.log
