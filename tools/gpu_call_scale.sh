python -m pytest tests/test_gpu_e2e.py -x -q -k "schedule" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "rc=$?"
tail -c 600 gpurun_out/r2_bench_2gpu.err
