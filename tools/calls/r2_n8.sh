# 8 GPUs: driver-shaped bench at N = 8 (and N = 4), multi-device parity test on 4 ranks
export SPICE_PREBUILT=1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -1 gpurun_out/r2_bench_n8.json | cut -c1-3600; tail -3 gpurun_out/r2_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --no-generation > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
tail -1 gpurun_out/r2_bench_n4.json | cut -c1-2600
timeout 600 python -m pytest tests/test_gpu_multi_device.py -m gpu -x -q 2>&1 | tail -2
