python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2j_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2j_smoke.log
python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err; echo "ref rc=$?"
timeout 600 python tools/config_profile.py C3/4 C4/2 C5 > gpurun_out/config_profile_mid.txt 2>&1; grep -E "bs|score_select" gpurun_out/config_profile_mid.txt
