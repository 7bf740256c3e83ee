"""Host-to-host inference helper: the call sites of the reference move `inputs['pose2d'].cuda()` in and
`pred_mesh.cpu()` out around every forward (lib/core/base.py:216,223-229).  For large batches the 83 KB/mesh
device-to-host copy is as long as a good part of the compute, so this helper cuts the batch into slices and
overlaps the copy of slice i with the forward of slice i+1 on a second stream (pinned host buffers)."""
from __future__ import annotations

import torch


def plan_slices(batch: int, ratio: float, floor: int = 148):
    """Slice boundaries [0, ..., batch]: sizes s, s r, s r^2, ... (each >= `floor` = one CTA per SM) that sum to the batch,
    so that each slice's device-to-host copy is shorter than the next slice's kernels and only a short last copy is exposed."""
    if batch <= 2 * floor:
        return [0, batch]
    sizes, nxt, left = [], float(batch) * (1.0 - ratio), batch
    while left > 0:
        sz = min(left, max(floor, int(nxt)))
        if left - sz < floor:
            sz = left
        sizes.append(sz)
        left -= sz
        nxt *= ratio
    bounds = [0]
    for sz in sizes:
        bounds.append(bounds[-1] + sz)
    return bounds


class HostPipeline:
    """forward(pose2d_host (B,J,2) pinned) -> (mesh_host (B,6890,3), pose3d_host (B,J,3)) pinned, both fp32.
    The call returns after the device-to-host copies have completed (it synchronises its copy stream), so the results can
    be read immediately; the two pinned buffers belong to the pipeline and are overwritten by the next forward()."""

    RATIO = 0.35

    def __init__(self, model, batch: int, slice_samples: int = 0):
        """slice_samples = 0: automatic.  The copy of slice i is hidden behind the kernels of slice i+1 as long as it is
        shorter than them, and only the last slice's copy is exposed, so the slices shrink geometrically by RATIO, a
        bound on (copy time per mesh) / (compute time per mesh): 83 KB at PCIe 5 x16 ~ 1.5 us against ~5 us of kernels
        on a B200 (measured 56 GB/s; a time-based calibration was tried and dropped - one noisy sample costs more than
        the fixed margin does)."""
        p = next(model.parameters())
        self.model, self.dev = model, p.device
        self.J = model.num_joint
        self.batch = batch
        self.slice_samples = slice_samples
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.mesh_host = torch.empty((batch, 6890, 3), dtype=torch.float32).pin_memory()
        self.pose3d_host = torch.empty((batch, self.J, 3), dtype=torch.float32).pin_memory()
        self._keep = []
        self._bounds = None

    def _plan(self):
        return plan_slices(self.batch, self.RATIO)

    @torch.no_grad()
    def forward(self, pose2d_host: torch.Tensor):
        B = pose2d_host.shape[0]
        if B != self.batch:
            raise ValueError(f'pipeline was built for batch {self.batch}, got {B}')
        if B == 0:
            return self.mesh_host, self.pose3d_host
        main = torch.cuda.current_stream(self.dev)
        m = self.model
        if self.slice_samples:
            bounds = list(range(0, B, self.slice_samples)) + [B]
        else:
            if self._bounds is None:
                self._bounds = self._plan()
            bounds = self._bounds
        self._keep.clear()
        # the lifter runs once over the whole batch (its outputs are small: 12 J + 512 J bytes per sample) ...
        xd = pose2d_host.to(self.dev, non_blocking=True)
        p3, feat = m.pose_lifter(xd.reshape(B, self.J * 2))
        p3 = p3.reshape(B, self.J, 3)
        done = torch.cuda.Event()
        done.record(main)
        self.copy_stream.wait_event(done)
        with torch.cuda.stream(self.copy_stream):
            self.pose3d_host.copy_(p3, non_blocking=True)
        p3.record_stream(self.copy_stream)
        # ... the decoder slice by slice, each slice's 83 KB/mesh copy overlapping the next slice's kernels
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            mesh = m.pose2mesh.forward_parts(xd[lo:hi], p3[lo:hi], feat[lo:hi])
            done = torch.cuda.Event()
            done.record(main)
            self.copy_stream.wait_event(done)
            with torch.cuda.stream(self.copy_stream):
                self.mesh_host[lo:hi].copy_(mesh, non_blocking=True)
            mesh.record_stream(self.copy_stream)
            self._keep.append(mesh)
        main.wait_stream(self.copy_stream)
        # host-to-host contract: the returned pinned buffers are complete when this call returns (the caller reads them
        # with numpy straight away, base.py:223-229).  They are reused by the next forward(): consume or copy them first.
        self.copy_stream.synchronize()
        self._keep.clear()
        return self.mesh_host, self.pose3d_host
