import torch, sys
sys.path.insert(0, '/root/repo')
import polyblur_b200 as pb
from polyblur_b200 import synthetic
x = synthetic.make("mosaic", 4, 3, 2160, 3840, device="cuda")
for _ in range(2): pb.domain_transform.recursive_filter(x, 2.0, 0.8, 1)
x2 = synthetic.make("mosaic", 4, 3, 1080, 1920, device="cuda")
for _ in range(2): pb.domain_transform.normalized_convolution(x2, 60.0, 0.4, 3)
for _ in range(2): pb.filters.bilateral_filter(x)
torch.cuda.synchronize()
