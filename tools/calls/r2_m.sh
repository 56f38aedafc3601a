export SPICE_PREBUILT=1
timeout 900 python -m pytest tests/test_gpu_generator.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_generator.py 10000 31623 100000 | tee gpurun_out/r2m_generator.jsonl | cut -c1-330
SPICE_GEN_TIMING=1 timeout 200 python tools/c2_sharded.py --n 1000000 --check-blocks 4 --block-rows 16 2> gpurun_out/r2m_c2_timing.err | cut -c1-900
grep "1000000 x" gpurun_out/r2m_c2_timing.err | awk '{k=$6; for(i=7;i<NF-1;i++) k=k" "$i; a[k]+=$(NF-1); n[k]++} END{for(k in a) print k, a[k], n[k]}'
