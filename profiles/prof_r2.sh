# Round-2 capture at the DEFAULT bench command (so per-launch numbers match the bench line).  The launch list is
# filtered to the kernels of the fit / vote / draw steps: process_dataset (outside the metric) alone launches
# thousands of rank-coding kernels (their list: profiles/r1f_encode_launches.csv).
set -x
CMD="python bench.py --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/plainF.json 2> gpurun_out/plainF.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"grow_kernel|slot_gather_kernel|predict_kernel|pack_kernel|encode_kernel|leaves_kernel|FillFunctor" -c 400 --csv --log-file gpurun_out/launchesF.csv $CMD > gpurun_out/ncu_launchF.log 2>&1
echo rc_launch=$?
FAST="python bench.py --no-cpu-baseline --e2e-steps 1 --no-c2a --c4-rows 0 --c5-rows 0"
ncu --set full --clock-control none --import-source on -k regex:grow_kernel -s 3 -c 1 -o gpurun_out/prof_growF $FAST > gpurun_out/ncu_growF.log 2>&1
echo rc_grow=$?
ncu --set full --clock-control none --import-source on -k regex:slot_gather_kernel -s 3 -c 1 -o gpurun_out/prof_gatherF $FAST > gpurun_out/ncu_gatherF.log 2>&1
echo rc_gather=$?
ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 3 -c 1 -o gpurun_out/prof_predictF $FAST > gpurun_out/ncu_predF.log 2>&1
echo rc_pred=$?
