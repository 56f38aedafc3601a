export SPICE_PREBUILT=1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
B="--steps 20 --warmup 5 --no-cpu-baseline --no-generation"
timeout 600 python bench.py $B > gpurun_out/r2r_pipe.json 2> gpurun_out/r2r_pipe.err; tail -2 gpurun_out/r2r_pipe.err
SPICE_PIPELINE=0 timeout 600 python bench.py $B > gpurun_out/r2r_nopipe.json 2> gpurun_out/r2r_nopipe.err
for f in gpurun_out/r2r_pipe.json gpurun_out/r2r_nopipe.json; do grep -o '"value[^,]*\|"ms_per_step[^,]*\|"frac[^,]*\|update_ms_total[^,]*\|deliver_ms_total[^,]*\|matches_reference"[^,]*\|"windows[^,]*' $f | tr '\n' ' '; echo; done
