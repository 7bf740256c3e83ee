"""Host-to-host inference helper: the call sites of the reference move `inputs['pose2d'].cuda()` in and
`pred_mesh.cpu()` out around every forward (lib/core/base.py:216,223-229).  For large batches the 83 KB/mesh
device-to-host copy is as long as a good part of the compute, so this helper cuts the batch into slices and
overlaps the copy of slice i with the forward of slice i+1 on a second stream (pinned host buffers)."""
from __future__ import annotations

import torch


def plan_slices(batch: int, ratio: float, floor: int = 296):
    """Slice boundaries [0, ..., batch]: sizes s, s r, s r^2, ... (each >= `floor` samples: the persistent kernels run two CTAs per SM) that sum to the batch,
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

    RATIO = 0.6

    def __init__(self, model, batch: int, slice_samples: int = 0):
        """slice_samples = 0: automatic.  The copy of slice i is hidden behind the kernels of slice i+1 as long as it is
        shorter than them, and only the last slice's copy is exposed, so the slices shrink geometrically by RATIO, a
        bound on (copy time per mesh) / (decoder time per mesh): 83 KB at PCIe 5 x16 = 1.5 us (measured 56 GB/s) against
        1.7 us of decoder kernels on a B200 (round 2; it was 5 us in round 1, when RATIO was 0.35).  Every slice also
        costs ~0.18 ms of fixed decoder time (20 launches, partial waves), which is why the ratio is not higher: measured
        at B = 4096 (tools/pipeline_sweep.py) 0.35 -> 13.0 ms, 0.5 -> 12.0, 0.6 -> 11.45, 0.7 -> 11.5, 0.8 -> 11.8-12.0,
        0.9 -> 12.8 ms per step against 9.04 ms of kernels and 6.16 ms of copy.  A time-based calibration was tried and
        dropped - one noisy sample costs more than the fixed margin does.  For throughput use submit() / result():
        whole-batch forwards whose copies overlap the NEXT batch's kernels."""
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
        self._ring = None                   # submit()/result(): two pinned output sets, allocated on first use
        self._tickets = 0

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

    # ---- throughput mode: double-buffered across batches --------------------------------------------------------------
    @torch.no_grad()
    def submit(self, pose2d_host: torch.Tensor) -> int:
        """Queue one batch and return a ticket without waiting: H2D of the inputs, the whole-batch forward (no slicing:
        best kernel efficiency) and the D2H of mesh + pose3d on the copy stream, which runs while the NEXT submit()'s
        kernels execute.  Two pinned output sets alternate; result(ticket) must have been called for ticket - 2 before
        submit() reuses its buffers (enforced: submit() waits for that copy, and raises if it was never collected)."""
        B = pose2d_host.shape[0]
        if B != self.batch:
            raise ValueError(f'pipeline was built for batch {self.batch}, got {B}')
        if self._ring is None:
            self._ring = [dict(mesh=self.mesh_host, pose3d=self.pose3d_host, done=None, keep=None, ticket=-1, collected=True),
                          dict(mesh=torch.empty_like(self.mesh_host).pin_memory(),
                               pose3d=torch.empty_like(self.pose3d_host).pin_memory(), done=None, keep=None, ticket=-1,
                               collected=True)]
        slot = self._ring[self._tickets & 1]
        if not slot['collected']:
            raise RuntimeError(f'HostPipeline.submit: the result of ticket {slot["ticket"]} was never collected; its '
                               'buffers would be overwritten')
        main = torch.cuda.current_stream(self.dev)
        xd = pose2d_host.to(self.dev, non_blocking=True)
        mesh, p3 = self.model(xd)
        ready = torch.cuda.Event()
        ready.record(main)
        self.copy_stream.wait_event(ready)
        with torch.cuda.stream(self.copy_stream):
            slot['pose3d'].copy_(p3.reshape(B, self.J, 3), non_blocking=True)
            slot['mesh'].copy_(mesh, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        mesh.record_stream(self.copy_stream)
        p3.record_stream(self.copy_stream)
        slot.update(done=done, keep=(mesh, p3, xd), ticket=self._tickets, collected=False)
        self._tickets += 1
        return slot['ticket']

    def result(self, ticket: int):
        """Block until the batch of `ticket` is complete in pinned host memory; returns (mesh_host, pose3d_host) views
        that stay valid until the second submit() after this ticket."""
        slot = self._ring[ticket & 1] if self._ring is not None else None
        if slot is None or slot['ticket'] != ticket:
            raise ValueError(f'HostPipeline.result: ticket {ticket} is not outstanding')
        slot['done'].synchronize()
        slot['keep'] = None
        slot['collected'] = True
        return slot['mesh'], slot['pose3d']
