# delivery kernel variants on the bench network (N = 1): bulk copies vs cp.async, 8 vs 16 warps per CTA
export SPICE_PREBUILT=1
B="--steps 20 --warmup 5 --no-parity --no-generation --no-cpu-baseline --no-e2e"
timeout 600 python bench.py $B > gpurun_out/r2v_bulk8.json 2> gpurun_out/r2v_bulk8.err
SPICE_DELIVER_WARPS=16 timeout 600 python bench.py $B > gpurun_out/r2v_bulk16.json 2> gpurun_out/r2v_bulk16.err
SPICE_DELIVER_PATH=1 timeout 600 python bench.py $B > gpurun_out/r2v_cpa8.json 2> gpurun_out/r2v_cpa8.err
SPICE_DELIVER_PATH=1 SPICE_DELIVER_WARPS=16 timeout 600 python bench.py $B > gpurun_out/r2v_cpa16.json 2> gpurun_out/r2v_cpa16.err
SPICE_DELIVER_PATH=1 timeout 600 python -m pytest tests/test_gpu_sim.py -m gpu -q -x 2>&1 | tail -3
grep -h -o '"roofline".*"windows": [0-9]*' gpurun_out/r2v_*.json | cut -c1-330
