cd $GRAFT_REPO_ROOT
NWCWT_STREAMS=1 NWCWT_RING_MB=200 timeout 300 python bench.py --steps 3 --warmup 3 2>&1 | tail -c 1500
