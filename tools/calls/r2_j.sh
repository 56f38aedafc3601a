export SPICE_PREBUILT=1
timeout 900 python -m pytest tests/test_gpu_sim.py -m gpu -x -q 2>&1 | tail -2
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-generation"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'update|deliver|prologue|publish|wait|sink' -s 150 -c 400 --csv --log-file gpurun_out/r2_launches_j.csv $B > gpurun_out/r2j_ncu_l.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline --no-e2e --no-generation | grep -o '"ms_per_step[^,]*\|"frac[^,]*\|update_ms_total[^,]*\|deliver_ms_total[^,]*'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:update_stateful -s 60 -c 2 -o gpurun_out/r2j_update $B > gpurun_out/r2j_ncu_u.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:update_stateless -s 30 -c 1 -o gpurun_out/r2j_update_p $B > gpurun_out/r2j_ncu_p.log 2>&1
