# round 2: which mapping on the bench's own c2 / c3 maps (softmax-mean token maps): QUAD, LANE, drift thresholds
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
  echo "== $1"
  env $1 timeout 300 python bench.py --workload $2 --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/r02u_tmp.json 2>/dev/null
  python - <<'PY'
import json
d = json.load(open("gpurun_out/r02u_tmp.json"))
print("  value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "sustained", round(d["sustained"]["ms_per_step"], 4), {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in d["kernels"].items()})
PY
}
for wl in c2 c3; do
run "X=0" $wl
run "ATTWARP_QUAD_MAP=1" $wl
run "ATTWARP_QUAD_MAP=2" $wl
run "ATTWARP_QUAD_DRIFT=3" $wl
run "ATTWARP_QUAD_DRIFT=4" $wl
run "ATTWARP_QUAD_DRIFT=6" $wl
done
