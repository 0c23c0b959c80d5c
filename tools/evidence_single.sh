#!/bin/bash
# One-GPU evidence of a build, run under gpurun from the repo root:   bash tools/evidence_single.sh <tag>
# Everything lands in gpurun_out/<tag>/ ; tools/make_profiles.py turns the .ncu-rep files into the committed summaries.
# Every ncu command runs only after the same command exited 0 without ncu (B200_PROFILING.md).
TAG=${1:-evidence}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -q -m gpu > $OUT/pytest_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1
python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_n1.json 2> $OUT/bench_reference_n1.err
python bench.py --steps 2 --warmup 3 --no-per-config --no-cpu > $OUT/bench_plain.json 2> $OUT/bench_plain.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench_c5.csv \
      python bench.py --steps 2 --warmup 3 --no-per-config --no-cpu > $OUT/ncu_launch.log 2>&1
for wl in c5 c4 c2 c3; do
  python tools/prof_run.py $wl 4 > $OUT/prof_$wl.txt 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:hvb_gen -s 2 -c 1 -o $OUT/${TAG}_gen_$wl \
        python tools/prof_run.py $wl 4 > $OUT/ncu_$wl.log 2>&1
done
tail -3 $OUT/pytest_gpu.txt; tail -4 $OUT/smoke.txt; cat $OUT/bench_n1.json | head -c 1500; echo; cat $OUT/bench_reference_n1.json | head -c 600; echo
for wl in c5 c4 c2 c3; do tail -1 $OUT/prof_$wl.txt; done
ls -la $OUT
