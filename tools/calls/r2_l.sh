export SPICE_PREBUILT=1
timeout 900 python -m pytest tests/test_gpu_generator.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_generator.py 10000 31623 100000 | tee gpurun_out/r2l_generator.jsonl | cut -c1-330
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r2l_launches_gen.csv python tools/bench_generator.py 100000 > gpurun_out/r2l_ncu_g.log 2>&1
timeout 200 python tools/c2_sharded.py --n 200000 --check-blocks 16 --block-rows 128 | cut -c1-900
SPICE_GEN_FORCE_EXACT=1 timeout 200 python tools/c2_sharded.py --n 100000 --check-blocks 8 --block-rows 64 | cut -c1-500
