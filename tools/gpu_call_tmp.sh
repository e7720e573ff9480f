timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "score or select or importance" > gpurun_out/cl_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/cl_tests.log
timeout 300 python tools/score_small_bench.py > gpurun_out/score_small_cluster.txt 2>&1; cat gpurun_out/score_small_cluster.txt
