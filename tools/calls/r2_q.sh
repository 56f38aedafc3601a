export SPICE_PREBUILT=1
timeout 900 python -m pytest tests/test_gpu_sim.py tests/test_gpu_samples.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/bench_samples.py --steps 3000 > gpurun_out/r2q_samples.jsonl 2> gpurun_out/r2q_samples.err; python - <<'PY'
import json
for l in open("gpurun_out/r2q_samples.jsonl"):
    d=json.loads(l); print(d["network"], {k:(round(v["ms_per_step"],5) if isinstance(v,dict) and "ms_per_step" in v else v) for k,v in d.items() if k.startswith("gpu") or k.startswith("cpu") or k.startswith("speed")})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 2100 --csv --log-file gpurun_out/r2q_launches_bp.csv python tools/brunel_plus_probe.py > gpurun_out/r2q.log 2>&1
