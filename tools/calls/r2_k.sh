export SPICE_PREBUILT=1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; cat gpurun_out/r2k_bench.json | grep -o '"value[^,]*\|"ms_per_step[^,]*\|"frac[^,]*\|update_ms_total[^,]*\|deliver_ms_total[^,]*\|matches_reference"[^,]*'
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-generation"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:update_stateful -s 60 -c 2 -o gpurun_out/r2k_update $B > gpurun_out/r2k_ncu_u.log 2>&1
timeout 300 python tools/bench_samples.py --steps 3000 > gpurun_out/r2k_samples.jsonl 2> gpurun_out/r2k_samples.err; cut -c1-600 gpurun_out/r2k_samples.jsonl
