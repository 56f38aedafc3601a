# leaner delivery kernel (zero-filled cp.async, absolute counter addresses, descriptors in shared memory), warp-per-row fp_rows_exact
export SPICE_PREBUILT=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_gputest.log 2>&1; echo rc=$? >> gpurun_out/r2c_gputest.log; tail -4 gpurun_out/r2c_gputest.log
B="--steps 20 --warmup 5 --no-parity --no-cpu-baseline --no-e2e"
timeout 600 python bench.py $B > gpurun_out/r2c_cpa8.json 2> gpurun_out/r2c_cpa8.err
SPICE_DELIVER_PATH=0 timeout 600 python bench.py $B --no-generation > gpurun_out/r2c_bulk8.json 2> gpurun_out/r2c_bulk8.err
SPICE_DELIVER_WARPS=16 timeout 600 python bench.py $B --no-generation > gpurun_out/r2c_cpa16.json 2> gpurun_out/r2c_cpa16.err
grep -h -o '"roofline".*"windows": [0-9]*' gpurun_out/r2c_cpa8.json gpurun_out/r2c_bulk8.json gpurun_out/r2c_cpa16.json | cut -c1-330
grep -h -o '"generation".*' gpurun_out/r2c_cpa8.json
N="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-generation"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deliver_units -s 40 -c 2 -o gpurun_out/r2c_deliver $N > gpurun_out/r2c_ncu_f.log 2>&1
timeout 300 python tools/bench_generator.py 10000 100000 > gpurun_out/r2c_generator.jsonl 2> gpurun_out/r2c_generator.err; cat gpurun_out/r2c_generator.jsonl
