python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests4.log 2>&1; echo "tests rc=$?"
python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref2.json 2> gpurun_out/r2_bench_ref2.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 204 -c 420 --csv --log-file gpurun_out/r2_step_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_step.log 2>&1; echo "ncu rc=$?"
python tools/kbench.py > gpurun_out/r2_kbench.log 2>&1; echo "kbench rc=$?"
