"""CUDA-graph capture of a whole optimisation step (forward + loss + backward).

The step issues ~600 launches of which most are O(rays)-sized torch ops of a few microseconds: replaying it as one
CUDA graph removes the launch gaps (the reference runs the same loop eagerly under autograd anomaly mode,
volsdf_train.py:20,556-578).  Shapes are static per iteration type (tracking / mapping), so one graph per type
suffices; inputs live in static device buffers that the caller overwrites (e.g. with cudaMemcpyAsync from pinned
host memory) before each replay.
"""
import torch


class GraphedStep:
    def __init__(self, fn, warmup=3):
        """fn() -> tensor (e.g. the loss); must read its inputs from tensors that stay alive (static buffers) and must
        not synchronise with the host."""
        # Warm-up and capture run on ONE side stream: autograd's AccumulateGrad nodes remember the stream of the first backward
        # pass that created them, and a capture on a different stream would fork into that stream for every parameter gradient
        # (torch warns "AccumulateGrad node's stream does not match ..."), leaving the accumulation on a branch of the graph.
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.output = fn()

    def replay(self):
        self.graph.replay()
        return self.output
