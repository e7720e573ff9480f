timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2h_tests.log
timeout 300 python tools/c2_small_batch_probe.py > gpurun_out/small_batch_pipe2.log 2>&1; cat gpurun_out/small_batch_pipe2.log
timeout 600 python tools/config_profile.py C3/8 C4/8 > gpurun_out/config_profile_shard2.txt 2>&1; grep -E "bs|attention" gpurun_out/config_profile_shard2.txt
