timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2e_tests.log
python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python tools/kbench.py > gpurun_out/r2e_kbench.log 2>&1; echo "kbench rc=$?"; grep attention gpurun_out/r2e_kbench.log
timeout 300 python tools/config_profile.py C1 C3 > gpurun_out/r2e_config_profile.txt 2>&1; grep -E "bs|attention" gpurun_out/r2e_config_profile.txt
