# 2 GPUs: the multi-device parity tests (real CUDA IPC / NVLink peer stores), then the driver-shaped bench at N = 2
export SPICE_PREBUILT=1
timeout 900 python -m pytest tests/test_gpu_multi_device.py -m gpu -x -q > gpurun_out/r2_gputest_n2.log 2>&1; echo rc=$? >> gpurun_out/r2_gputest_n2.log
tail -5 gpurun_out/r2_gputest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -2 gpurun_out/r2_bench_n2.json | cut -c1-3000; tail -5 gpurun_out/r2_bench_n2.err
SPICE_FLATTEN=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-parity --no-generation --no-cpu-baseline > gpurun_out/r2_bench_n2_flatten.json 2> gpurun_out/r2_bench_n2_flatten.err
tail -2 gpurun_out/r2_bench_n2_flatten.json | cut -c1-2000
