#!/bin/bash
# Round-end evidence run (1 GPU): GPU test suite, bench line, ncu launch list of the bench command, ncu --set full of one
# full train step.  Usage: bash scratch/final_profile.sh <tag>   (outputs under gpurun_out/<tag>_*)
tag=${1:-r2z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 26 --launch-count 13 -f -o gpurun_out/prof_${tag}_full \
    python profiles/profile_step.py cfg3 4 full > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
tail -c 600 gpurun_out/${tag}_bench.json
tail -2 gpurun_out/${tag}_ncu_full.log
