L=universal-beta-splatting_b200/ubs_b200/lib
(time python -m pytest tests -m gpu -x -q) 2>&1 | tail -6
cp $L/libubs_b200.so /tmp/new.so
for v in new rforig new rforig; do
if [ $v = new ]; then cp /tmp/new.so $L/libubs_b200.so; else cp $L/libv_rforig.so $L/libubs_b200.so; fi
echo "== $v"; python scratch/stage_bench.py cfg3 2>&1 | tail -1; python scratch/stage_bench.py cfg2 2>&1 | tail -1
done
cp /tmp/new.so $L/libubs_b200.so
python scratch/full_step_bench.py cfg3 2>&1 | tail -1
