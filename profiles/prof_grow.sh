set -x
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --predict-rows 100000"
$CMD > gpurun_out/plain2.json 2> gpurun_out/plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:grow_kernel -s 1 -c 1 -o gpurun_out/prof_grow2 $CMD > gpurun_out/ncu_grow2.log 2>&1
echo rc=$?
tail -3 gpurun_out/ncu_grow2.log
