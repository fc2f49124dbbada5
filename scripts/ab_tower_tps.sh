SZB_TOWER_TPS=2 timeout 400 python -m pytest tests/test_gpu_net.py tests/test_gpu_search.py -m gpu -x -q 2>&1 | tail -3
for t in 1 2 1 2; do
  SZB_TOWER_TPS=$t timeout 200 python bench.py --no-extras --steps 8 --warmup 3 2>/dev/null | grep "^{" > gpurun_out/ab_tps_$t.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_tps_$t.json").read()); r=d["roofline"]
print("tps", $t, round(d["value"]), r["ms_per_launch"], round(r["frac"],4), r.get("sm_mhz_inside_launches"), r.get("tensor_pipe_utilisation_at_that_clock"))
PY
done
