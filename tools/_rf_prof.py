import sys, torch
sys.path.insert(0, '/root/repo')
import polyblur_b200 as pb
from polyblur_b200 import synthetic
x = synthetic.make("mosaic", 4, 3, 2160, 3840, device="cuda")
for _ in range(2):
    y = pb.domain_transform.recursive_filter(x, 2.0, 0.8, 1)
    z = pb.filters.bilateral_filter(x)
torch.cuda.synchronize()
