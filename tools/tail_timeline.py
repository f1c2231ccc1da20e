"""Print the clock64() timeline of the bottleneck-tail kernels (VAD_TAIL_DEBUG=1): one I3Res50 forward at 160 clip-crops."""
import os
import sys

os.environ["VAD_TAIL_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anomaly_detection_on_video_b200.i3d import I3Res50

m = I3Res50().eval().cuda()
x = torch.randn(160, 16, 224, 232, 4, device="cuda").to(torch.bfloat16)
m.forward_stem_layout(x)
torch.cuda.synchronize()
