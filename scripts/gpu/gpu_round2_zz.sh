# round 2, last call: the reference's time stepper with the device pressure solve (drop-in) against the pure reference build
mkdir -p gpurun_out
timeout 25 python -m pytest tests/test_timestepper_dropin.py -x -q -m gpu > gpurun_out/r2zz_pytest.log 2>&1; tail -4 gpurun_out/r2zz_pytest.log
