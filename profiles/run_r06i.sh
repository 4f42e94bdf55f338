# round 2, pass r06i: ncu of the tensor-core LANCZOS kernel (64 x 24^2 -> 1344^2)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r06i
cat > /tmp/mota_drive.py <<'PY'
import torch, sys
sys.path.insert(0, ".")
from attwarp_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
t = torch.rand(64, 24, 24, device="cuda", generator=g)
for _ in range(4):
    m = ops.mota_mask(t, (1344, 1344))
    mx, my = ops.maps_from_mota_tokens(t, (1344, 1344))
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lanczos_up_mma -s 4 -c 2 -o ${P}_prof_lanczos_mma -f python /tmp/mota_drive.py > ${P}_ncu.log 2>&1; tail -n 3 ${P}_ncu.log
