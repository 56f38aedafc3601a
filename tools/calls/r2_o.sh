export SPICE_PREBUILT=1
timeout 900 python -m pytest tests -m gpu -q -rf 2>&1 | tail -8 | cut -c1-300
timeout 300 python tools/bench_adjlist.py | tee gpurun_out/r2o_adjlist.json | cut -c1-600
