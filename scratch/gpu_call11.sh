python bench.py --steps 50 --warmup 5 > gpurun_out/r1d_bench.json 2> gpurun_out/r1d_bench.err
tail -c 300 gpurun_out/r1d_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_b_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:rasterize_fwd|fused_project_bwd' --launch-skip 2 --launch-count 2 -f -o gpurun_out/prof_r1d_two python profiles/profile_step.py cfg3 3 full > gpurun_out/r1d_ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
