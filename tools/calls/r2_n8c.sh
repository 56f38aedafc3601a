export SPICE_PREBUILT=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/c2_sharded.py --size 1000000 --check-blocks 64 --block-rows 256 > gpurun_out/r2_c2_1e6_n8.json 2> gpurun_out/r2_c2_1e6_n8.err
cat gpurun_out/r2_c2_1e6_n8.json | cut -c1-1500; tail -3 gpurun_out/r2_c2_1e6_n8.err | cut -c1-300
