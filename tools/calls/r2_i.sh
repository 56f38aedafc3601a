export SPICE_PREBUILT=1
for i in 1 2 3; do timeout 900 python -m pytest tests/test_gpu_samples.py tests/test_gpu_sim.py -m gpu -q -rf 2>&1 | tail -15 | cut -c1-300; done
