timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" > gpurun_out/wi_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/wi_tests.log
timeout 300 python tools/attn_ab.py 121:87 152:121 100:70 197:173 > gpurun_out/attn_ab_warpitems.txt 2>&1; cat gpurun_out/attn_ab_warpitems.txt
timeout 300 python tools/attn_ab.py 256 16 143:128 128:115 92:82 > gpurun_out/attn_ab_warpitems_c4.txt 2>&1; cat gpurun_out/attn_ab_warpitems_c4.txt
