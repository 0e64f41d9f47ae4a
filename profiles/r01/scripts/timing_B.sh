# historical: the persistent pass-B kernel / NW_TIMING hooks this script measured were removed again (see shape_sweep.md)
cd $GRAFT_REPO_ROOT
NWCWT_LIB=$GRAFT_REPO_ROOT/ninwavelets_b200/libnwcwt_tim.so timeout 300 python bench.py --steps 1 --warmup 3 2>&1 | grep "passB cta" | sort | uniq -c | sort -rn | head -5
NWCWT_LIB=$GRAFT_REPO_ROOT/ninwavelets_b200/libnwcwt_tim.so timeout 300 python bench.py --steps 1 --warmup 3 2>&1 | grep "passB cta" | awk '{w+=$6; l+=$9; n++} END{print "avg wait", w/n, "avg lifetime", l/n, "n", n, "frac", w/l}'
