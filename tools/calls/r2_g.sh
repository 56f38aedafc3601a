export SPICE_PREBUILT=1
timeout 900 python -m pytest tests/test_gpu_sim.py -m gpu -x -q 2>&1 | tail -2
B="--steps 20 --warmup 5 --no-parity --no-cpu-baseline --no-e2e --no-generation"
timeout 600 python bench.py $B > gpurun_out/r2g_bulk8.json 2> gpurun_out/r2g_bulk8.err
grep -h -o '"roofline".*"windows": [0-9]*' gpurun_out/r2g_bulk8.json | cut -c1-230
N="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-generation"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deliver_units -s 40 -c 1 -o gpurun_out/r2g_deliver_bulk $N > gpurun_out/r2g_ncu_fb.log 2>&1
