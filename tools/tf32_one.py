"""Two I3Res50 forwards in the TF32 precision mode at B clip-crops (ncu target: skip the first forward's launches)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomaly_detection_on_video_b200.engine import ingest_ncthw_tf32
from anomaly_detection_on_video_b200.i3d import I3Res50
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
m = I3Res50().eval().to(dev)
m.precision = "tf32"
plan = m.plan(dev)
xs = ingest_ncthw_tf32(torch.randn(B, 3, 16, 224, 224, device=dev), planes=bool(plan.ops[0].flags & 64))
for _ in range(2):
    f = plan.forward(xs)
torch.cuda.synchronize()
print("ok", tuple(f.shape), [op.name for op in m.op_table()][:60])
