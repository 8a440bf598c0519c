# Round-2b capture (second half of round 2: split-bitmap grower, small-tree vote kernel).  Every program runs once
# WITHOUT ncu first (its line is the bench value); the ncu passes are shares / per-launch counters only.  The
# .ncu-rep files are turned into raw-page CSVs on the box and deleted (gpurun brings back at most 64 MiB).
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python bench.py > gpurun_out/bench_r2b_n1.json 2> gpurun_out/bench_r2b_n1.err; echo rc_bench=$?
CMD="python bench.py --no-cpu-baseline --e2e-steps 1"
cp gpurun_out/bench_r2b_n1.json gpurun_out/plainF.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"grow_kernel|slot_gather_kernel|predict_kernel|predict_small_kernel|pack_kernel|encode_kernel|leaves_kernel|FillFunctor" -c 400 --csv --log-file gpurun_out/launchesF.csv $CMD > gpurun_out/ncu_launchF.log 2>&1
echo rc_launch=$?
FAST="python bench.py --no-cpu-baseline --e2e-steps 1 --no-c2a --c4-rows 0 --c5-rows 0"
ncu --set full --clock-control none --import-source on -k regex:grow_kernel -s 3 -c 1 -o gpurun_out/prof_growF $FAST > gpurun_out/ncu_growF.log 2>&1
echo rc_grow=$?
# the small-tree vote kernel on the C5 forest, 1,048,576 rows of the shard (the same launch shape per CTA as 12.5M rows)
C5="python bench.py --no-cpu-baseline --e2e-steps 1 --no-c2a --c4-rows 0 --c5-rows 1048576"
ncu --set full --clock-control none --import-source on -k regex:"predict_small_kernel" -s 2 -c 1 -o gpurun_out/prof_smallF $C5 > gpurun_out/ncu_smallF.log 2>&1
echo rc_small=$?
for r in growF smallF; do
  ncu -i gpurun_out/prof_$r.ncu-rep --page raw --csv > gpurun_out/prof_$r.raw.csv 2> /dev/null
  ncu -i gpurun_out/prof_$r.ncu-rep --page source --csv 2> /dev/null | head -c 3000000 > gpurun_out/prof_$r.source.csv
done
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
