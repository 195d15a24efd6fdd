"""Tracking-loop specialisation: the 50-100 pose iterations of one frame
(/root/reference/code/training/volsdf_train.py:394-446) as ONE captured CUDA graph replayed per iteration.

An iteration = draw the pixel set (when ``change_pixels``, volsdf_train.py:41-43,411-412), gather their ground truth from
the device-resident FrameCache, forward in ``mode="tracking"`` (pose-only: no grid scatter, no weight gradients), tracking
loss, backward to the 7-vector pose, Adam step (+ StepLR(50, 0.95) as a device-side scale), best-candidate bookkeeping
(``if loss < current_min_loss``, :436-438) -- everything on the device, no host round trip inside the loop; the reference
runs the same loop eagerly at R = 1024 rays, where it is launch-latency bound.
"""
import torch

from .utils.general import get_camera_from_tensor


class TrackingLoop:
    def __init__(self, model, loss, cache, num_pixels=1024, lr=1e-3, change_pixels=True, lr_step=50, lr_gamma=0.95, use_graph=True):
        self.model, self.loss, self.cache = model, loss, cache
        self.num_pixels, self.base_lr, self.change_pixels = int(num_pixels), float(lr), bool(change_pixels)
        self.lr_step, self.lr_gamma = int(lr_step), float(lr_gamma)
        dev = cache.device
        self.cam = torch.zeros(7, device=dev, requires_grad=True)
        self.slot = torch.zeros(1, dtype=torch.long, device=dev)
        self.sidx = torch.zeros(self.num_pixels, dtype=torch.long, device=dev)
        self.best_loss = torch.full((), 1e10, device=dev)
        self.best_cam = torch.zeros(7, device=dev)
        self.first_loss = torch.zeros((), device=dev)
        self.last_loss = torch.zeros((), device=dev)
        self.it = torch.zeros((), dtype=torch.long, device=dev)
        # Adam(lr) on the pose with the learning rate as a device tensor so that the StepLR decay lives inside the graph
        self.lr_t = torch.tensor(self.base_lr, device=dev)
        self.opt = torch.optim.Adam([self.cam], lr=self.lr_t, capturable=dev.type == "cuda", foreach=False)
        self.frame_idx = 0
        self.graph = None
        self.use_graph = use_graph and dev.type == "cuda"

    # one pose iteration on the static buffers
    def _iteration(self):
        if self.change_pixels:
            self.sidx.copy_(torch.randint(self.cache.total_pixels, (self.num_pixels,), device=self.sidx.device))
        idx, inp, gt = self.cache.batch(None, self.sidx, slots=self.slot)
        inp = dict(inp)
        inp["pose"] = get_camera_from_tensor(self.cam).unsqueeze(0)
        out = self.model(inp, self.slot, gt, mode="tracking", frame_idx=self.frame_idx)
        loss = self.loss(out, gt, stage="fine", frame_idx=self.frame_idx)["loss"]
        self.opt.zero_grad(set_to_none=False)
        loss.backward()
        with torch.no_grad():
            better = loss.detach() < self.best_loss                      # candidate = the pose that PRODUCED this loss (:436-438)
            self.best_cam.copy_(torch.where(better, self.cam.detach(), self.best_cam))
            self.best_loss.copy_(torch.where(better, loss.detach(), self.best_loss))
            self.first_loss.copy_(torch.where(self.it == 0, loss.detach(), self.first_loss))
            self.last_loss.copy_(loss.detach())
        self.opt.step()
        with torch.no_grad():
            self.it += 1
            self.lr_t.copy_(self.base_lr * self.lr_gamma ** torch.div(self.it, self.lr_step, rounding_mode="floor").float())
        return loss

    def _reset(self, cam7_init, frame_idx):
        with torch.no_grad():
            self.cam.copy_(cam7_init.to(self.cam.device))
            self.slot.copy_(self.cache.slots([frame_idx]))
            self.best_loss.fill_(1e10)
            self.best_cam.copy_(self.cam)
            self.it.zero_()
            self.lr_t.fill_(self.base_lr)
            for st in self.opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
            if not self.change_pixels:
                self.sidx.copy_(torch.randint(self.cache.total_pixels, (self.num_pixels,), device=self.sidx.device))
        self.frame_idx = int(frame_idx)

    def track(self, frame_idx, cam7_init, iters):
        """Optimise the pose of cached frame ``frame_idx`` from ``cam7_init`` (quat wxyz + translation) for ``iters``
        iterations.  Returns (best cam7, first loss, last loss) as device tensors (no sync)."""
        was_training = self.model.training
        self.model.train()
        self._reset(cam7_init, frame_idx)
        if self.use_graph:
            if self.graph is None or self._graph_frame_is_zero != (self.frame_idx == 0):
                # frame_idx only enters the loss through "frame_idx == 0" switches (loss.py:179-184): one graph per case
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for _ in range(3):
                        self._iteration()
                torch.cuda.current_stream().wait_stream(s)
                self._reset(cam7_init, frame_idx)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=s):      # same stream as the warm-up (see graph.py)
                    self._iteration()
                self._graph_frame_is_zero = self.frame_idx == 0
                self._reset(cam7_init, frame_idx)
            for _ in range(iters):
                self.graph.replay()
        else:
            for _ in range(iters):
                self._iteration()
        self.model.train(was_training)
        return self.best_cam.clone(), self.first_loss.clone(), self.last_loss.clone()
