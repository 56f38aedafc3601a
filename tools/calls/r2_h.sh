export SPICE_PREBUILT=1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "--- rounds forced on the test networks"
SPICE_PREZEROED=1 SPICE_DELIVER_ROUNDS=3 timeout 900 python -m pytest tests/test_gpu_sim.py tests/test_gpu_samples.py -m gpu -x -q 2>&1 | tail -2
echo "--- rank shapes on one GPU"
for G in 8 4 2; do
  echo "G=$G whole units (16 warps)"; SPICE_PREZEROED=0 timeout 300 python tools/rank_shape_probe.py $G 300 0.5 | tail -1 | cut -c1-400
  echo "G=$G rounds auto"; SPICE_PREZEROED=1 timeout 300 python tools/rank_shape_probe.py $G 300 0.5 | tail -1 | cut -c1-400
done
for R in 1 2 3 4 6; do echo "G=8 rounds=$R"; SPICE_PREZEROED=1 SPICE_DELIVER_ROUNDS=$R timeout 300 python tools/rank_shape_probe.py 8 300 0.5 | tail -1 | cut -c1-400; done
echo "--- N=1 bench with consumer-zeroed counters"
SPICE_PREZEROED=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline --no-e2e --no-generation | grep -o '"ms_per_step[^,]*\|"frac[^,]*\|update_ms_total[^,]*'
timeout 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline --no-e2e --no-generation | grep -o '"ms_per_step[^,]*\|"frac[^,]*\|update_ms_total[^,]*'
echo "--- c2 sharded tool, small, one rank"
timeout 300 python tools/c2_sharded.py --n 20000 --check-blocks 8 --block-rows 64 | cut -c1-700
SPICE_GEN_TIMING=1 timeout 300 python tools/bench_generator.py 100000 2>/dev/null | tail -1
