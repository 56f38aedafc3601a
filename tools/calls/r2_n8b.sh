# 8 GPUs: BASELINE configs[1] at 1e6 x 1e6 sharded over the ranks (checked against the oracle), then the driver-shaped bench at N = 8
export SPICE_PREBUILT=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/c2_sharded.py --n 1000000 --check-blocks 64 --block-rows 256 > gpurun_out/r2_c2_1e6_n8.json 2> gpurun_out/r2_c2_1e6_n8.err
cat gpurun_out/r2_c2_1e6_n8.json | cut -c1-1500; tail -3 gpurun_out/r2_c2_1e6_n8.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-generation > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err
tail -1 gpurun_out/r2_bench_n8b.json | cut -c1-3000; tail -3 gpurun_out/r2_bench_n8b.err | cut -c1-300
