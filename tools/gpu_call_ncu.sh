set -x
python tools/attn_one.py 197 197 > gpurun_out/plain_attn_r2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 -o gpurun_out/r2_attn_pipe_197 -f python tools/attn_one.py 197 197 > gpurun_out/ncu_attn_r2.log 2>&1
python tools/attn_one.py 197 173 > gpurun_out/plain_attn_r2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 -o gpurun_out/r2_attn_pipe_173 -f python tools/attn_one.py 197 173 > gpurun_out/ncu_attn_r2b.log 2>&1
python tools/gemm_one.py fc1 > gpurun_out/plain_fc1_r2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/r2_gemm_fc1 -f python tools/gemm_one.py fc1 > gpurun_out/ncu_fc1_r2.log 2>&1
python tools/gemm_one.py proj > gpurun_out/plain_proj_r2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/r2_gemm_proj -f python tools/gemm_one.py proj > gpurun_out/ncu_proj_r2.log 2>&1
python tools/gemm_one.py fc2 > gpurun_out/plain_fc2_r2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/r2_gemm_fc2 -f python tools/gemm_one.py fc2 > gpurun_out/ncu_fc2_r2.log 2>&1
ls -la gpurun_out/*.ncu-rep
