"""Host-to-host inference helper: the call sites of the reference move `inputs['pose2d'].cuda()` in and
`pred_mesh.cpu()` out around every forward (lib/core/base.py:216,223-229).  For large batches the 83 KB/mesh
device-to-host copy is as long as a good part of the compute, so this helper cuts the batch into slices and
overlaps the copy of slice i with the forward of slice i+1 on a second stream (pinned host buffers)."""
from __future__ import annotations

import torch


class HostPipeline:
    """forward(pose2d_host (B,J,2) pinned) -> (mesh_host (B,6890,3), pose3d_host (B,J,3)) pinned, both fp32."""

    def __init__(self, model, batch: int, slices: int = 4):
        p = next(model.parameters())
        self.model, self.dev = model, p.device
        self.J = model.num_joint
        self.batch = batch
        self.slices = max(1, min(slices, batch))
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.mesh_host = torch.empty((batch, 6890, 3), dtype=torch.float32).pin_memory()
        self.pose3d_host = torch.empty((batch, self.J, 3), dtype=torch.float32).pin_memory()
        self._keep = []

    @torch.no_grad()
    def forward(self, pose2d_host: torch.Tensor):
        B = pose2d_host.shape[0]
        if B != self.batch:
            raise ValueError(f'pipeline was built for batch {self.batch}, got {B}')
        main = torch.cuda.current_stream(self.dev)
        step = (B + self.slices - 1) // self.slices
        self._keep.clear()
        for lo in range(0, B, step):
            hi = min(B, lo + step)
            xd = pose2d_host[lo:hi].to(self.dev, non_blocking=True)
            mesh, p3 = self.model(xd)
            done = torch.cuda.Event()
            done.record(main)
            self.copy_stream.wait_event(done)
            with torch.cuda.stream(self.copy_stream):
                self.mesh_host[lo:hi].copy_(mesh, non_blocking=True)
                self.pose3d_host[lo:hi].copy_(p3, non_blocking=True)
            mesh.record_stream(self.copy_stream)
            p3.record_stream(self.copy_stream)
            self._keep.append((mesh, p3))
        main.wait_stream(self.copy_stream)
        return self.mesh_host, self.pose3d_host
