(time python -m pytest tests -m gpu -x -q) 2>&1 | tail -5
for c in cfg3 cfg2 cfg4 cfg3_r4; do python scratch/stage_bench.py $c 2>&1 | tail -1; done
python bench.py --steps 50 --warmup 5 > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err
tail -c 300 gpurun_out/r1e_bench.err
