export SPICE_PREBUILT=1
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-generation"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_b.csv $B > gpurun_out/r2_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deliver_units -s 40 -c 2 -o gpurun_out/r2_deliver_b $B > gpurun_out/r2_ncu_f.log 2>&1
timeout 600 python tools/bench_samples.py --steps 3000 > gpurun_out/r2_samples.jsonl 2> gpurun_out/r2_samples.err
timeout 300 python tools/bench_generator.py 10000 31623 100000 > gpurun_out/r2_generator.jsonl 2> gpurun_out/r2_generator.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_gen.csv python tools/bench_generator.py 100000 > gpurun_out/r2_ncu_g.log 2>&1
tail -3 gpurun_out/r2_samples.jsonl gpurun_out/r2_generator.jsonl; tail -3 gpurun_out/r2_samples.err gpurun_out/r2_ncu_f.log
