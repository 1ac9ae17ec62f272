export PYTHONDONTWRITEBYTECODE=1
sel=dwconv,outlook
for tag in base v2; do OGV_LIB=$PWD/outlook_grid_vision_transformer_b200/_ab/libogvit_$tag.so python tools/kbench.py --only $sel --stages 0,1,2,3 --reps 20 --out gpurun_out/kb_$tag.json > /dev/null 2>&1; done
python tools/kbench.py --only $sel --stages 0,1,2,3 --reps 20 --out gpurun_out/kb_new.json > /dev/null 2>&1
python - <<PY
import json
L={t:{r["kernel"]:r for r in json.load(open(f"gpurun_out/kb_{t}.json"))} for t in ("base","v2","new")}
for k in L["base"]:
    b=L["base"][k]["ms"]
    print(f"{k:24s} base {b*1e3:8.1f}  v2 {L['v2'][k]['ms']*1e3:8.1f} ({100*(L['v2'][k]['ms']/b-1):+5.1f}%)  new {L['new'][k]['ms']*1e3:8.1f} ({100*(L['new'][k]['ms']/b-1):+5.1f}%)  {100*L['new'][k]['frac']:5.1f}% HBM")
PY
