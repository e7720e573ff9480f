timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" > gpurun_out/qke_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/qke_tests.log
timeout 300 python tools/attn_ab.py 197:197 197:173 173:152 152:152 87:87 121:87 152:121 > gpurun_out/attn_ab_qkempty.txt 2>&1; cat gpurun_out/attn_ab_qkempty.txt
timeout 300 python tools/attn_ab.py 32 12 197:197 197:173 121:87 > gpurun_out/attn_ab_qkempty_b32.txt 2>&1; cat gpurun_out/attn_ab_qkempty_b32.txt
