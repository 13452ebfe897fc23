mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_r2Z.log 2>&1; tail -2 gpurun_out/pytest_r2Z.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2Z.log 2>&1; tail -2 gpurun_out/smoke_r2Z.log
timeout 900 python bench.py > gpurun_out/bench_r2Z.json 2> gpurun_out/bench_r2Z.err; tail -c 300 gpurun_out/bench_r2Z.err; tail -1 gpurun_out/bench_r2Z.json | cut -c1-200
timeout 600 python tools/c5_full_check.py gpurun_out/c5_full_r2Z.json > gpurun_out/c5_full_r2Z.log 2>&1; tail -1 gpurun_out/c5_full_r2Z.log | cut -c1-400
