# scratch script for one gpurun call (overwritten per call during development); this is the round's last validation call
python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python tools/kbench.py > gpurun_out/kbench.log 2>&1; grep attention gpurun_out/kbench.log
