import torch, sys, json
sys.path.insert(0, '/root/repo')
import polyblur_b200 as pb
from polyblur_b200 import synthetic, _lib
x = synthetic.make("mosaic", 4, 3, 2160, 3840, device="cuda")
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("rf 1 iter 4x4K ms", timed(lambda: pb.domain_transform.recursive_filter(x, 2.0, 0.8, 1)))
print("rf 3 iter 4x4K ms", timed(lambda: pb.domain_transform.recursive_filter(x, 60.0, 0.4, 3)))
x2 = synthetic.make("mosaic", 4, 3, 1080, 1920, device="cuda")
print("nc 3 iter 4x1080p ms", timed(lambda: pb.domain_transform.normalized_convolution(x2, 60.0, 0.4, 3)))
print("polyblur rf prefilter 4x4K n3 ms", timed(lambda: pb.polyblur_deblurring(x, n_iter=3, alpha=6, beta=1, prefiltering=True, prefilter="rf")))
print("polyblur no prefilter 4x4K n3 ms", timed(lambda: pb.polyblur_deblurring(x, n_iter=3, alpha=6, beta=1)))
print("bilateral 4x4K ms", timed(lambda: pb.filters.bilateral_filter(x)))
_lib.profile_begin(); pb.domain_transform.recursive_filter(x, 2.0, 0.8, 1); torch.cuda.synchronize(); print(_lib.profile_end())
