cd tools/probes && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../rajni_vit_b200/csrc gather4_rate_probe.cu -o /tmp/gather4_rate_probe -lcuda 2>/dev/null; cd ../..
timeout 60 /tmp/gather4_rate_probe > gpurun_out/gather4_rate_probe.txt 2>&1; echo "rc=$?"; cat gpurun_out/gather4_rate_probe.txt
