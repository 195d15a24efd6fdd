"""Timing + accuracy of nicer_outer_accum on the bench shapes (M=64, N in {64, 71, 3+...}, P = 401408 / 90112)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nicer_slam_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
for P in (401408, 90112):
    for M, N in ((64, 64), (64, 71), (64, 33), (3, 64), (1, 64)):
        A = torch.randn(M, P, device=dev)
        B = torch.randn(N, P, device=dev)
        Cm = torch.zeros(M, N, device=dev)
        b = torch.zeros(M, device=dev)
        ops.outer_accum(A, B, Cm, b)
        ref = (A.double() @ B.double().t())
        err = float((Cm.double() - ref).norm() / ref.norm())
        berr = float((b.double() - A.double().sum(1)).norm() / A.double().sum(1).norm())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            ops.outer_accum(A, B, Cm, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        gb = (M + N) * P * 4 / 1e9
        print(f"P={P} M={M} N={N}: {ms*1e3:7.1f} us  {gb/ms*1e3/1e3:5.2f} TB/s  rel err {err:.2e} bias {berr:.2e}")

