export SPICE_PREBUILT=1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
grep -h -o '"roofline".*"windows": [0-9]*' gpurun_out/r2f_bench.json | cut -c1-230
cut -c1-400 gpurun_out/r2f_bench.json
N="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-generation"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deliver_units -s 40 -c 1 -o gpurun_out/r2f_deliver_bulk $N > gpurun_out/r2f_ncu_fb.log 2>&1
