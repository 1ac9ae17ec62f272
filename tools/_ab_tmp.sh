export PYTHONDONTWRITEBYTECODE=1
sel=colsum,colstats,bn_bwd,se_pool,mbconv_bwd_stats,dw_bn2,rowscale
OGV_LIB=$PWD/outlook_grid_vision_transformer_b200/_ab/libogvit_base.so python tools/kbench.py --only $sel --stages 0,1,2,3 --reps 30 --out gpurun_out/kb_base.json > /dev/null 2>&1
python tools/kbench.py --only $sel --stages 0,1,2,3 --reps 30 --out gpurun_out/kb_new.json > /dev/null 2>&1
python - <<PY
import json
b={r["kernel"]:r["ms"] for r in json.load(open("gpurun_out/kb_base.json"))}
n={r["kernel"]:r for r in json.load(open("gpurun_out/kb_new.json"))}
for k in b: print(f"{k:30s} base {b[k]*1e3:8.1f}  new {n[k]['ms']*1e3:8.1f}  {100*(n[k]['ms']/b[k]-1):+6.1f}%   {100*n[k]['frac']:5.1f}% HBM")
PY
