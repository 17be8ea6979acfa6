cd $GRAFT_REPO_ROOT
timeout 300 python scripts/e2e_timeline.py --out gpurun_out/r02_tl2.json > /dev/null 2> gpurun_out/r02_tl2.err
rm -f gpurun_out/*_chrome.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_tl2.json"))
print("wall", d["wall_ms"], "span", d["gpu_span_ms"])
for k, v in d["rows"].items():
    if k.startswith("memcpy"): print(k, v)
for c in d["copies_over_1mb"]: print(c)
PY
tail -3 gpurun_out/r02_tl2.err
