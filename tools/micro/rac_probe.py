import sys; sys.path.insert(0, '.')
import torch
from vlnce_b200 import obs_transforms as ot
g = torch.Generator().manual_seed(6)
f = {"rgb": torch.randint(0, 256, (64, 480, 640, 3), generator=g, dtype=torch.uint8).cuda(),
     "depth": torch.rand(64, 480, 640, 1, generator=g).cuda()}
tr = [ot.ResizeShortestEdge(256), ot.CenterCropperPerSensor([("rgb", (224, 224)), ("depth", (256, 256))])]
for _ in range(3):
    ot.apply_obs_transforms_batch(dict(f), tr)
torch.cuda.synchronize()
