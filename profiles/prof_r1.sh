set -x
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --n-trees 64 --predict-rows 200000"
$CMD > gpurun_out/plain.json 2> gpurun_out/plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo rc_launch=$?
ncu --set full --clock-control none --import-source on -k regex:grow_kernel -s 1 -c 1 -o gpurun_out/prof_grow $CMD > gpurun_out/ncu_grow.log 2>&1
echo rc_grow=$?
ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 1 -c 1 -o gpurun_out/prof_predict $CMD > gpurun_out/ncu_pred.log 2>&1
echo rc_pred=$?
cat gpurun_out/plain.json | cut -c1-1500
tail -3 gpurun_out/ncu_grow.log
