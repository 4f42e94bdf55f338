cd $GRAFT_REPO_ROOT
for g in 1 0; do
for cap in 2 3; do
for st in 1 2 3 4; do
ATTWARP_REMAP_CTAS_PER_SM=$cap python profiles/overlap_probe.py --streams $st --sets $((st*2)) --graph $g --steps 120
done; done; done
