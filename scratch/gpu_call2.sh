python -m pytest tests/test_gpu_fused.py tests/test_gpu_train_step.py -x -q -m gpu 2>&1 | tail -5
for cfg in cfg3 cfg2 cfg4; do
for pair in "5 4" "6 5" "8 6" "7 4"; do
set -- $pair
echo "== $cfg FWD_MINB=$1 BWD_MINB=$2"
UBS_FWD_MINB=$1 UBS_BWD_MINB=$2 python scratch/stage_bench.py $cfg bwd 2>&1 | tail -1
done
done
