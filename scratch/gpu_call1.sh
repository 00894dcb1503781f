set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r1c_pytest.log 2>&1
tail -3 gpurun_out/r1c_pytest.log
python bench.py > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1c_bench_ref.json 2> gpurun_out/r1c_bench_ref.err
python bench.py --impl reference_cuda --steps 20 --warmup 3 > gpurun_out/r1c_bench_refcuda.json 2> gpurun_out/r1c_bench_refcuda.err
python scratch/counts.py cfg3 > gpurun_out/r1c_counts.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_b_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:rasterize|fused_project|bin_|segment_sort|ssim|loss_finalize|adam' --launch-skip 11 --launch-count 11 -f -o gpurun_out/prof_r1c_full python profiles/profile_step.py cfg3 3 full > gpurun_out/r1c_ncu_full.log 2>&1
ls -la gpurun_out
