# 2 GPUs: NCCL parity test of the final code, then the e2e fit with each rank held to 4 host cores (what a rank
# has on the 8-GPU box: 32 cores / 8 ranks) -- host walk of numpy's stream against the device walk.
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dist.py -x -q -m gpu -p no:cacheprovider > gpurun_out/n2_dist_test.log 2>&1; echo rc_dist=$?; tail -3 gpurun_out/n2_dist_test.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
Q="bench.py --gpus 2 --no-c2a --c4-rows 0 --c5-rows 0 --no-cpu-baseline --steps 3 --warmup 3"
taskset -c 0-7 $RUN $Q --host-streams 4 > gpurun_out/n2_host4.json 2> gpurun_out/n2_host4.err; echo rc_host4=$?
RFMC_DEVICE_DRAWS=1 taskset -c 0-7 $RUN $Q --host-streams 4 > gpurun_out/n2_dev4.json 2> gpurun_out/n2_dev4.err; echo rc_dev4=$?
RFMC_DEVICE_DRAWS=1 taskset -c 0-7 $RUN $Q --host-streams 8 > gpurun_out/n2_dev8.json 2> gpurun_out/n2_dev8.err; echo rc_dev8=$?
grep -h "e2e step" gpurun_out/n2_host4.err | tail -3; grep -h "e2e step" gpurun_out/n2_dev4.err | tail -3; grep -h "e2e step" gpurun_out/n2_dev8.err | tail -3
