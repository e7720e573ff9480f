timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" > gpurun_out/rot_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/rot_tests.log
timeout 300 python tools/attn_ab.py 152:152 173:152 173:173 197:173 138:138 159:143 > gpurun_out/attn_ab_rot3.txt 2>&1; cat gpurun_out/attn_ab_rot3.txt
RAJNI_ATTN_ROT=1 timeout 300 python tools/attn_ab.py 152:152 173:152 173:173 197:173 138:138 159:143 > gpurun_out/attn_ab_rot1.txt 2>&1; cat gpurun_out/attn_ab_rot1.txt
RAJNI_ATTN_ROT=4 timeout 300 python tools/attn_ab.py 152:152 173:152 138:138 > gpurun_out/attn_ab_rot4.txt 2>&1; cat gpurun_out/attn_ab_rot4.txt
RAJNI_ATTN_ROT=2 timeout 300 python tools/attn_ab.py 152:152 173:152 173:173 197:197 > gpurun_out/attn_ab_rot2.txt 2>&1; cat gpurun_out/attn_ab_rot2.txt
