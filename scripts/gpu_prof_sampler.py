"""ncu target: the sampler's sdf-only pass (coarse + fine) on bench-shaped inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nicer_slam_b200 import ops

dev = torch.device("cuda:0")
step = bench.build_step(device=dev)
m = step.model
U = step.rays * bench.N_EVAL
x = (torch.rand(U, 3, device=dev) * 2 - 1) * 0.9
nets = [m.implicit_network.coarse.fused_args(), m.implicit_network.fine.fused_args()]
for _ in range(3):
    ops.sdf_values(x, nets)
torch.cuda.synchronize()
