set -x
cd $GRAFT_REPO_ROOT
export SPICE_PREBUILT=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02b_smi.txt
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r02b_pytest.log 2>&1
tail -15 gpurun_out/r02b_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -c 2500 gpurun_out/r02b_bench.json
SPICE_DELIVER_WARPS=16 timeout 600 python bench.py --steps 20 --warmup 5 --no-parity --no-generation --no-cpu-baseline --no-e2e > gpurun_out/r02b_bench_w16.json 2> gpurun_out/r02b_bench_w16.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deliver_units -s 20 -c 1 -o gpurun_out/r02b_deliver python bench.py --steps 1 --warmup 0 --time-steps 15 --no-e2e --no-parity --no-generation --no-cpu-baseline > gpurun_out/r02b_ncu_bench.json 2> gpurun_out/r02b_ncu.err
ls -la gpurun_out
