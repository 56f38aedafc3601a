set -x
cd $GRAFT_REPO_ROOT
export SPICE_PREBUILT=1
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-parity --no-generation --no-cpu-baseline --no-e2e"
(time timeout 900 python -m pytest tests/test_gpu_sim.py tests/test_gpu_samples.py -m gpu -q -x) > gpurun_out/r02c_pytest.log 2>&1
tail -5 gpurun_out/r02c_pytest.log
(time SPICE_DELIVER_PATH=1 timeout 900 python -m pytest tests/test_gpu_sim.py -m gpu -q -x) > gpurun_out/r02c_pytest_path1.log 2>&1
tail -5 gpurun_out/r02c_pytest_path1.log
timeout 600 python bench.py $B > gpurun_out/r02c_bench_bulk8.json 2> gpurun_out/r02c_bench_bulk8.err
SPICE_DELIVER_WARPS=16 timeout 600 python bench.py $B > gpurun_out/r02c_bench_bulk16.json 2> gpurun_out/r02c_bench_bulk16.err
SPICE_DELIVER_PATH=1 timeout 600 python bench.py $B > gpurun_out/r02c_bench_cpa8.json 2> gpurun_out/r02c_bench_cpa8.err
SPICE_DELIVER_PATH=1 SPICE_DELIVER_WARPS=16 timeout 600 python bench.py $B > gpurun_out/r02c_bench_cpa16.json 2> gpurun_out/r02c_bench_cpa16.err
grep -h -o '"roofline".*"windows": [0-9]*' gpurun_out/r02c_bench_*.json | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deliver_units -s 20 -c 1 -o gpurun_out/r02c_deliver python bench.py --steps 1 --warmup 0 --time-steps 15 --no-e2e --no-parity --no-generation --no-cpu-baseline > gpurun_out/r02c_ncu_bench.json 2> gpurun_out/r02c_ncu.err
SPICE_DELIVER_PATH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:deliver_units -s 20 -c 1 -o gpurun_out/r02c_deliver_cpa python bench.py --steps 1 --warmup 0 --time-steps 15 --no-e2e --no-parity --no-generation --no-cpu-baseline > gpurun_out/r02c_ncu_bench_cpa.json 2> gpurun_out/r02c_ncu_cpa.err
ls -la gpurun_out | tail -20
