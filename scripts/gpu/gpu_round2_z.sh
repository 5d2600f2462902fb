# round 2, call z (1 GPU, the last seconds): the GPU test on right-hand sides captured from the reference's own time steps
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "captured_time_step" > gpurun_out/r2z_pytest.log 2>&1; tail -3 gpurun_out/r2z_pytest.log
