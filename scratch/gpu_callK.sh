python bench.py --steps 50 --warmup 5 > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err
tail -c 300 gpurun_out/r1f_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1f_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_b_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:rasterize|fused_project|bin_|segment_sort|ssim|loss_finalize|adam' --launch-skip 11 --launch-count 11 -f -o gpurun_out/prof_r1f_full python profiles/profile_step.py cfg3 3 full > gpurun_out/r1f_ncu_full.log 2>&1
python bench.py --impl reference_cuda --steps 20 --warmup 3 > gpurun_out/r1f_bench_refcuda.json 2>/dev/null
ls -la gpurun_out | tail -8
