"""Occupancy-grid fill during NeRF training: the other consumer of the density
kernel.  ``train_ngp_nerf.py:262-337`` calls ``occupancy_grid.every_n_step(step, occ_eval_fn)`` with
``occ_eval_fn(x) = query_density(x) * render_step_size`` (:267-290); the grid is nerfacc 0.3.5's
``OccupancyGrid`` (un-vendored; restated from its published behaviour, PARITY UNPINNED):

  * every ``n`` steps: during warm-up all cells, afterwards ``num_cells / 4`` uniformly drawn cells plus as many
    drawn from the currently occupied ones;
  * one jittered point per chosen cell, mapped from the unit cube to the ROI;
  * ``occs[cell] = max(occs[cell] * ema_decay, occ)``; ``binary = occs > min(mean(occs), occ_thre)``.

The density of all chosen points comes from ONE call of the fused hash-grid + MLP kernel (``drb_ngp_density``);
``binary`` has the [R, R, R] bool layout that ``SampleGrid.set_binary_fields`` / ``extract_block`` consume.
"""
import torch
from torch import nn


class OccupancyGrid(nn.Module):
    NUM_DIM = 3

    def __init__(self, roi_aabb, resolution=128, contraction_type=None) -> None:
        super().__init__()
        if isinstance(resolution, (list, tuple)):
            assert len(set(resolution)) == 1, "cubic grids only"
            resolution = int(resolution[0])
        if not isinstance(roi_aabb, torch.Tensor):
            roi_aabb = torch.tensor(roi_aabb, dtype=torch.float32)
        self.res = int(resolution)
        self.num_cells = self.res ** 3
        self.register_buffer("_roi_aabb", roi_aabb.float())
        self.register_buffer("resolution", torch.tensor([self.res] * 3, dtype=torch.int32))
        self.register_buffer("occs", torch.zeros(self.num_cells))
        self.register_buffer("_binary", torch.zeros([self.res] * 3, dtype=torch.bool))

    @property
    def binary(self):
        return self._binary

    @torch.no_grad()
    def _cells_to_update(self, step, warmup_steps, generator):
        dev = self.occs.device
        if step < warmup_steps:
            return torch.arange(self.num_cells, device=dev)
        n = self.num_cells // 4
        uniform = torch.randint(self.num_cells, (n,), device=dev, generator=generator)
        occupied = torch.nonzero(self._binary.flatten())[:, 0]
        if n < occupied.numel():
            occupied = occupied[torch.randint(occupied.numel(), (n,), device=dev, generator=generator)]
        return torch.cat([uniform, occupied])

    @torch.no_grad()
    def _update(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, generator=None):
        idx = self._cells_to_update(step, warmup_steps, generator)
        r = self.res
        coords = torch.stack([idx // (r * r), (idx // r) % r, idx % r], dim=1).float()
        x = (coords + torch.rand(coords.shape, device=coords.device, generator=generator)) / r
        lo, hi = self._roi_aabb[:3], self._roi_aabb[3:]
        x = x * (hi - lo) + lo                                    # ContractionType.AABB inverse
        occ = occ_eval_fn(x).reshape(-1).to(self.occs.dtype)
        self.occs[idx] = torch.maximum(self.occs[idx] * ema_decay, occ)
        thre = torch.clamp(self.occs.mean(), max=occ_thre)
        self._binary = (self.occs > thre).reshape(r, r, r)

    @torch.no_grad()
    def every_n_step(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16, generator=None):
        """nerfacc OccupancyGrid.every_n_step (train_ngp_nerf.py:293)."""
        if not self.training:
            raise RuntimeError("every_n_step() is a training-time update; call .train() first (as nerfacc does)")
        if step % n == 0:
            self._update(step, occ_eval_fn, occ_thre, ema_decay, warmup_steps, generator)

    @torch.no_grad()
    def fill_from_field(self, radiance_field, render_step_size, step=0, **kw):
        """``every_n_step`` with the occ_eval_fn of train_ngp_nerf.py:267-290 (bounded scenes):
        density(x) * render_step_size, one fused-kernel call over all chosen cells."""
        self.every_n_step(step, lambda x: radiance_field.query_density(x) * render_step_size, **kw)
        return self._binary
