timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c_tests.log
