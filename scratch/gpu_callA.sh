L=universal-beta-splatting_b200/ubs_b200/lib
cp $L/libv_b11.so $L/libubs_b200.so
python -m pytest tests/test_gpu_bin_sort.py tests/test_gpu_golden.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3
for v in b11 b9 b11 b9; do
cp $L/libv_$v.so $L/libubs_b200.so
echo "== $v"; for c in cfg3 cfg2 cfg4 cfg3_r4; do python scratch/stage_bench.py $c 2>&1 | tail -1; done
done
cp $L/libv_b11.so $L/libubs_b200.so
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:segment_sort_kernel -s 8 -c 2 python scratch/stage_bench.py cfg3 2>&1 | grep -E "gpu__time|inst_executed" | tail -6
