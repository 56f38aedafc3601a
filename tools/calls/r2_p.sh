export SPICE_PREBUILT=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 2100 --csv --log-file gpurun_out/r2p_launches_bp.csv python tools/brunel_plus_probe.py > gpurun_out/r2p.log 2>&1
tail -2 gpurun_out/r2p.log
