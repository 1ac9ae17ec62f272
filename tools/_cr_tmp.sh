export PYTHONDONTWRITEBYTECODE=1
echo "== base lib"; OGV_LIB=$PWD/outlook_grid_vision_transformer_b200/_ab/libogvit_base.so python tools/kbench.py --only colsum,colstats,bn_bwd_reduce --stages 0,1,2,3 --reps 30 2>&1 | grep "^s[0-9]"
for cap in 0 592 148 74 37; do echo "== new lib cap=$cap"; OGV_CR_CTAS=$cap python tools/kbench.py --only colsum,colstats,bn_bwd_reduce --stages 0,1,2,3 --reps 30 2>&1 | grep "^s[0-9]"; done
