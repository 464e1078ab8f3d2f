"""Where an eager polyblur_deblurring(CUDA tensor) call spends its wall time: host enqueue, device work, sync."""
import os, sys, time, statistics
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import polyblur_b200
from polyblur_b200 import synthetic, deblurring

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = synthetic.make("mosaic", B, 3, 1080, 1920, device="cuda")
for _ in range(3):
    polyblur_b200.polyblur_deblurring(x, n_iter=3, alpha=6, beta=1)
torch.cuda.synchronize()
enq, wall, dev = [], [], []
for _ in range(9):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    y = polyblur_b200.polyblur_deblurring(x, n_iter=3, alpha=6, beta=1)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    enq.append((t1 - t0) * 1e3); wall.append((t2 - t0) * 1e3); dev.append(e0.elapsed_time(e1))
print("eager: host enqueue ms", round(statistics.median(enq), 3), "wall ms", round(statistics.median(wall), 3), "device (events) ms", round(statistics.median(dev), 3))
g = deblurring.GraphedPolyblur((B, 3, 1080, 1920), n_iter=3, alpha=6, beta=1)
g.x.copy_(x)
for _ in range(3):
    g.graph.replay()
torch.cuda.synchronize()
wall = []
for _ in range(9):
    torch.cuda.synchronize(); t0 = time.perf_counter(); g.graph.replay(); torch.cuda.synchronize(); wall.append((time.perf_counter() - t0) * 1e3)
print("graph replay: wall ms", round(statistics.median(wall), 3))
