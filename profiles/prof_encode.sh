# Launch list + one full capture of the dominant kernel of the device rank coding (process_dataset).
set -x
CMD="python profiles/encode_probe.py"
$CMD > gpurun_out/encode_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launchesE.csv $CMD > gpurun_out/ncu_launchE.log 2>&1
echo rc_launch=$?
ncu --set full --clock-control none --import-source on -k regex:bitonic_global_pass_kernel -s 40 -c 1 -o gpurun_out/prof_bitonicE $CMD > gpurun_out/ncu_bitonicE.log 2>&1
echo rc_bitonic=$?
ncu --set full --clock-control none --import-source on -k regex:interleave_kernel -c 1 -o gpurun_out/prof_interleaveE $CMD > gpurun_out/ncu_interleaveE.log 2>&1
echo rc_interleave=$?
cat gpurun_out/encode_plain.log
