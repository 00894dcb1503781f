for cfg in cfg3 cfg2; do
for pair in "5 4" "4 3" "3 2"; do
set -- $pair
echo "== $cfg FWD_MINB=$1 BWD_MINB=$2"
UBS_FWD_MINB=$1 UBS_BWD_MINB=$2 python scratch/stage_bench.py $cfg bwd 2>&1 | tail -1
done
done
