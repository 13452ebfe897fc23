mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2H.log 2>&1; tail -3 gpurun_out/smoke_r2H.log
python bench.py > gpurun_out/bench_r2H.json 2> gpurun_out/bench_r2H.err; tail -c 300 gpurun_out/bench_r2H.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2H.json 2>> gpurun_out/bench_r2H.err
python tools/prof_bvh.py > gpurun_out/prof_bvh_r2H.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'trace_rays_bvh|render_regen' -o gpurun_out/bvh_r2H python tools/prof_bvh.py > gpurun_out/ncu_bvh_r2H.log 2>&1
ls -la gpurun_out/bvh_r2H.ncu-rep
python tools/c5_full_check.py gpurun_out/c5_full_r2H.json > gpurun_out/c5_full_r2H.log 2>&1; tail -1 gpurun_out/c5_full_r2H.log | cut -c1-400
