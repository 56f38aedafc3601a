set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02a_pytest.log 2>&1
tail -5 gpurun_out/r02a_pytest.log
tools/build/bulk_run_bench 16 > gpurun_out/r02a_microbench.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 1500 gpurun_out/r02a_bench.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
ncu --set full --clock-control none --import-source on -k regex:deliver_tiles -s 20 -c 1 -o gpurun_out/r02a_deliver python bench.py --steps 1 --warmup 0 --time-steps 15 --no-e2e --no-parity --no-generation --no-cpu-baseline > gpurun_out/r02a_ncu_bench.json 2> gpurun_out/r02a_ncu.err
ls -la gpurun_out
