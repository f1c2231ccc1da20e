"""Smallest program that launches every kernel of the InceptionI3d op table once at 160 clip-crops (the ncu target of
tools/ncu_inception.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from anomaly_detection_on_video_b200.inception import InceptionI3d

dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = InceptionI3d()   # constructor initialisation: the kernels' time does not depend on the weights
m.eval().to(dev)
x = torch.randn(160, 16, 224, 232, 4, device=dev).to(torch.bfloat16)
f = m.forward_stem_layout(x)
torch.cuda.synchronize()
print("ok", tuple(f.shape), float(f.abs().mean()))
