export SPICE_PREBUILT=1
nvidia-smi --query-gpu=memory.total --format=csv,noheader
SPICE_GEN_TIMING=1 timeout 400 python tools/c2_sharded.py --n 1000000 --check-blocks 1 --block-rows 16 --emulate 3/8 2> gpurun_out/r2n_c2_timing.err | cut -c1-1200
tail -3 gpurun_out/r2n_c2_timing.err | cut -c1-300
grep "1000000 x" gpurun_out/r2n_c2_timing.err | awk '{k=$6; for(i=7;i<NF-1;i++) k=k" "$i; a[k]+=$(NF-1); n[k]++} END{for(k in a) print k, a[k], n[k]}'
