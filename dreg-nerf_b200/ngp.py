"""Host-side mirrors of the extract interfaces: ``NGPradianceField`` (conerf/radiance_fields/
ngp.py:66-193) and ``SampleGrid`` (conerf/register/sample_grid.py:59-343), plus the
``Evaluator.sample_points`` scatter (eval_ngp_nerf.py:337-412) as ``extract_block``.

The arithmetic runs in libdregb200 (fused hash-grid + MLP kernels, occupancy-grid ray marcher).
tiny-cuda-nn and nerfacc are NOT required.  Parameter layout follows tiny-cuda-nn's flat ``params``
tensors so that a reference ``model.pth`` loads: ``mlp_base.params`` = [W1 64x32 | W2 16x64 | hash
table], ``color_mlp.params`` = [C1 64x32 | C2 64x64 | C3 16x64], ``direction_encoding.params`` empty
(layout restated from tiny-cuda-nn's documentation; parity with the wheel is unpinned).
"""
import ctypes as C
import math
import struct

import torch
from torch import nn

from . import _lib

N_W1, N_W2 = 64 * 32, 16 * 64
N_C1, N_C2, N_C3 = 64 * 32, 64 * 64, 16 * 64


def hash_table_entries():
    """Entries ([2] floats each) of the 16-level table (levels concatenated): levels whose res^3 (rounded up to
    8) fits 2^19 are dense, the others hashed (tiny-cuda-nn grid encoding; same arithmetic as
    drb_ngp_table_entries, against which it is checked on first device use).  Pure Python so that building a
    field - e.g. for the CPU reference arm of bench.py - does not load the CUDA library."""
    total = 0
    for level in range(16):
        scale = struct.unpack("f", struct.pack("f", 2.0 ** (level * math.log2(1.4472692012786865)) * 16 - 1.0))[0]
        res = int(math.ceil(scale)) + 1
        total += min((res ** 3 + 7) // 8 * 8, 1 << 19)
    return total


def check_march_options(meta_data, cut_off):
    """The marcher uses nerfacc's fixed step (cone_angle == 0) and never drops samples by alpha
    (alpha_thre == 0): both are what eval_ngp_nerf.py:79-92 passes for bounded scenes.  A checkpoint trained
    with other settings would get different sample positions - refuse instead of producing a wrong mask
    (sample_grid.py:294-295, confidence_loss.py:137-138 forward both to ray_marching)."""
    cone = meta_data.get("cone_angle", 0.0) if hasattr(meta_data, "get") else 0.0
    thre = meta_data.get("alpha_thre", 0.0) if hasattr(meta_data, "get") else 0.0
    if cone not in (None, 0, 0.0):
        raise NotImplementedError("cone_angle=%r: libdregb200's marcher implements the fixed-step march only" % (cone,))
    if thre not in (None, 0, 0.0) and float(thre) > float(cut_off):
        raise NotImplementedError("alpha_thre=%r above the cut-off %r would drop decisive samples; not implemented"
                                  % (thre, cut_off))


class _FlatParams(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.params = nn.Parameter(torch.zeros(n, dtype=torch.float32))


class NGPradianceField(nn.Module):
    """Instant-NGP field with the reference's constructor and query surface."""

    def __init__(self, aabb, num_dim: int = 3, use_viewdirs: bool = True, density_activation=None,
                 unbounded: bool = False, geo_feat_dim: int = 15, n_levels: int = 16,
                 log2_hashmap_size: int = 19) -> None:
        super().__init__()
        if unbounded:
            raise NotImplementedError("unbounded (contracted) scenes are outside the registration hot path")
        if (num_dim, geo_feat_dim, n_levels, log2_hashmap_size, use_viewdirs) != (3, 15, 16, 19, True):
            raise NotImplementedError("kernels are specialised for the reference's configuration "
                                      "(3-D, 16 levels, T=2^19, 15 geometry features, view directions)")
        if not isinstance(aabb, torch.Tensor):
            aabb = torch.tensor(aabb, dtype=torch.float32)
        self.register_buffer("aabb", aabb.float())
        self.num_dim, self.geo_feat_dim, self.unbounded, self.use_viewdirs = 3, 15, False, True
        self.table_entries = hash_table_entries()
        self.mlp_base = _FlatParams(N_W1 + N_W2 + 2 * self.table_entries)
        self.direction_encoding = _FlatParams(0)
        self.color_mlp = _FlatParams(N_C1 + N_C2 + N_C3)
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self, table_std=None):
        """tiny-cuda-nn style init: hash table U(-1e-4, 1e-4) (or N(0, table_std) for synthetic fields
        whose density should cross the 0.7 threshold), MLPs xavier-uniform."""
        p = self.mlp_base.params
        for off, (rows, cols) in ((0, (64, 32)), (N_W1, (16, 64))):
            bound = math.sqrt(6.0 / (rows + cols))
            p[off:off + rows * cols].uniform_(-bound, bound)
        if table_std is None:
            p[N_W1 + N_W2:].uniform_(-1e-4, 1e-4)
        else:
            p[N_W1 + N_W2:].normal_(0.0, table_std)
        c = self.color_mlp.params
        off = 0
        for rows, cols in ((64, 32), (64, 64), (16, 64)):
            bound = math.sqrt(6.0 / (rows + cols))
            c[off:off + rows * cols].uniform_(-bound, bound)
            off += rows * cols

    def _params_struct(self):
        p, c = self.mlp_base.params, self.color_mlp.params
        if not p.is_cuda:
            raise _lib.DrbError("libdregb200 has no CPU path: move the field to a CUDA device")
        if self.table_entries != int(_lib.load().drb_ngp_table_entries()):
            raise _lib.DrbError("hash-table layout of the Python mirror and libdregb200 differ")
        base, cb = p.data_ptr(), c.data_ptr()
        s = _lib.NgpParams(hash_table=base + 4 * (N_W1 + N_W2), w1=base, w2=base + 4 * N_W1,
                           c1=cb, c2=cb + 4 * N_C1, c3=cb + 4 * (N_C1 + N_C2))
        for i, v in enumerate(self.aabb.detach().cpu().tolist()):
            s.aabb[i] = v
        return s

    @torch.no_grad()
    def query_density(self, x, return_feat: bool = False):
        """ngp.py:148-176: density [..., 1] (and geometry features [..., 15])."""
        lib = _lib.load()
        shape = list(x.shape[:-1])
        xf = x.reshape(-1, 3).contiguous().float()
        n = xf.shape[0]
        density = torch.empty(n, dtype=torch.float32, device=x.device)
        feat = torch.empty((n, 15), dtype=torch.float32, device=x.device) if return_feat else None
        ps = self._params_struct()
        with torch.cuda.device(x.device):
            _lib.check(lib.drb_ngp_density(C.byref(ps), _lib.ptr(xf), n, _lib.ptr(density), _lib.ptr(feat),
                                           _lib.stream_ptr()), "drb_ngp_density")
        density = density.reshape(shape + [1]).to(x)
        if return_feat:
            return density, feat.reshape(shape + [15]).to(x)
        return density

    @torch.no_grad()
    def query_rgb_mean(self, viewdirs, embedding):
        """Mean colour over a fixed table of view directions (sample_grid.py:332-340)."""
        lib = _lib.load()
        e = embedding.reshape(-1, 15).contiguous().float()
        n = e.shape[0]
        dirs = viewdirs.detach().cpu().float().contiguous()
        nd = dirs.shape[0]
        host = (C.c_float * (3 * nd))(*dirs.reshape(-1).tolist())
        rgb = torch.empty((n, 3), dtype=torch.float32, device=e.device)
        ps = self._params_struct()
        with torch.cuda.device(e.device):
            _lib.check(lib.drb_ngp_rgb_mean(C.byref(ps), _lib.ptr(e), n, host, nd, _lib.ptr(rgb),
                                            _lib.stream_ptr()), "drb_ngp_rgb_mean")
        return rgb

    @torch.no_grad()
    def query_rgb(self, dir, embedding):
        """ngp.py:178-193 for the call pattern of the extract path: one direction repeated for all
        points (sample_grid.py:333).  Per-point directions belong to NeRF rendering, which is out of
        scope for this library."""
        d = dir.reshape(-1, 3)
        if d.shape[0] > 1 and not bool((d == d[:1]).all()):
            raise NotImplementedError("per-point view directions (NeRF rendering) are out of scope")
        rgb = self.query_rgb_mean(d[:1], embedding)
        return rgb.reshape(list(embedding.shape[:-1]) + [3]).to(embedding)


def fixed_viewing_directions():
    """sample_grid.py:132-146, including its quirks: x == y and directions are not unit length."""
    dirs = []
    for phi in (math.pi / 3, 0, -math.pi):
        for k in range(6):
            theta = k * math.pi / 3
            dirs.append([math.cos(phi) * math.sin(theta), math.cos(phi) * math.sin(theta), math.sin(theta)])
    return torch.tensor(dirs, dtype=torch.float32)


class SampleGrid(nn.Module):
    """sample_grid.py:59-343 (AABB contraction only)."""
    NUM_DIM = 3

    def __init__(self, roi_aabb, resolution=128, contraction_type=None) -> None:
        super().__init__()
        if isinstance(resolution, (list, tuple)):
            assert len(set(resolution)) == 1, "cubic grids only"
            resolution = int(resolution[0])
        if isinstance(resolution, torch.Tensor):
            resolution = int(resolution.reshape(-1)[0])
        if not isinstance(roi_aabb, torch.Tensor):
            roi_aabb = torch.tensor(roi_aabb, dtype=torch.float32)
        assert roi_aabb.shape == torch.Size([6])
        self.res = int(resolution)
        self.num_voxels = self.res ** 3
        self.register_buffer("_binary", torch.zeros([self.res] * 3, dtype=torch.bool))
        self.register_buffer("resolution", torch.tensor([self.res] * 3, dtype=torch.int32))
        self.register_buffer("_roi_aabb", roi_aabb.float())
        self._delta = 1e-2
        self._viewdirs = fixed_viewing_directions()

    @property
    def binary(self):
        return self._binary

    @torch.no_grad()
    def set_binary_fields(self, binary):
        self._binary = binary

    @torch.no_grad()
    def uniform_sample_occupied_voxels(self):
        return torch.nonzero(self._binary.flatten())[:, 0]

    @torch.no_grad()
    def query_radiance_and_density_from_camera(self, radiance_field, occupancy_grid, meta_data, device,
                                               density_thre=0.7, cut_off: float = 0.5, jitter=None,
                                               return_grid: bool = False, surface_only_where_dense: bool = False,
                                               rgb_only_where_masked: bool = False, workspace=None):
        """sample_grid.py:208-343 -> (points, color, alpha, indices, density_mask, surface_mask).

        ``jitter`` (U[0,1) [K,3]) may be supplied for reproducibility; by default it is drawn with
        torch.rand on the grid's device exactly where the reference calls torch.rand_like.
        ``surface_only_where_dense`` skips the ray march of cells whose density fails the threshold
        (their surface_mask entries stay False): identical voxel_grid / voxel_mask, ~3x less marching.
        ``rgb_only_where_masked`` evaluates the colour head only for cells that pass both masks (their rows are
        the only ones voxel_grid keeps); the colour of every other cell is returned as 0.
        ``workspace``: a uint8 CUDA tensor of at least ``extract_workspace_bytes(k)`` bytes - the call then allocates
        nothing (drb_extract_block_ws); by default scratch comes from the stream-ordered allocator.
        """
        lib = _lib.load()
        check_march_options(meta_data, cut_off)
        device = torch.device(device)
        binary = self._binary.to(device)
        indices = torch.nonzero(binary.flatten())[:, 0].contiguous()
        k = indices.numel()
        if jitter is None:
            jitter = torch.rand((k, 3), dtype=torch.float32, device=device)
        jitter = jitter.to(device=device, dtype=torch.float32).contiguous()
        occ = occupancy_grid.binary if hasattr(occupancy_grid, "binary") else occupancy_grid
        occ_u8 = occ.to(device=device, dtype=torch.uint8).contiguous()
        cams = meta_data["camera_poses"].to(device=device, dtype=torch.float32)[:, :3, 3].contiguous()
        dirs = self._viewdirs.cpu().contiguous()
        host_dirs = (C.c_float * (3 * dirs.shape[0]))(*dirs.reshape(-1).tolist())
        desc = _lib.ExtractDesc(res=self.res, occupied=indices.data_ptr(), n_occupied=k,
                                jitter=jitter.data_ptr(), occ_binary=occ_u8.data_ptr(),
                                cam_origins=cams.data_ptr(), ncams=cams.shape[0],
                                render_step_size=float(meta_data["render_step_size"]),
                                density_thre=float(density_thre), cut_off=float(cut_off),
                                host_dirs=host_dirs, ndirs=dirs.shape[0],
                                surface_only_where_dense=int(surface_only_where_dense),
                                rgb_only_where_masked=int(rgb_only_where_masked))
        roi = self._roi_aabb.detach().cpu().tolist()
        scene = [float(v) for v in torch.as_tensor(meta_data["aabb"]).reshape(-1).tolist()]
        for i in range(6):
            desc.roi_aabb[i] = roi[i]
            desc.scene_aabb[i] = scene[i]
        f32 = dict(dtype=torch.float32, device=device)
        points, rgb = torch.empty((k, 3), **f32), torch.empty((k, 3), **f32)
        alpha = torch.empty((k, 1), **f32)
        dmask = torch.empty(k, dtype=torch.uint8, device=device)
        smask = torch.empty(k, dtype=torch.uint8, device=device)
        grid = torch.empty((self.res, self.res, self.res, 7), **f32) if return_grid else None
        ps = radiance_field._params_struct()
        with torch.cuda.device(device):
            if workspace is None:
                _lib.check(lib.drb_extract_block(C.byref(ps), C.byref(desc), _lib.ptr(points), _lib.ptr(rgb),
                                                 _lib.ptr(alpha), _lib.ptr(dmask), _lib.ptr(smask),
                                                 _lib.ptr(grid), _lib.stream_ptr()), "drb_extract_block")
            else:
                _lib.check(lib.drb_extract_block_ws(C.byref(ps), C.byref(desc), _lib.ptr(points), _lib.ptr(rgb),
                                                    _lib.ptr(alpha), _lib.ptr(dmask), _lib.ptr(smask), _lib.ptr(grid),
                                                    _lib.ptr(workspace), workspace.numel() * workspace.element_size(),
                                                    _lib.stream_ptr()), "drb_extract_block_ws")
            _lib.check_stream_flag("drb_extract_block")      # the caller indexes with the masks next: no extra stall
        out = (points, rgb, alpha, indices, dmask.bool(), smask.bool())
        return out + (grid,) if return_grid else out


def extract_workspace_bytes(n_occupied):
    """Bytes of scratch one extract of ``n_occupied`` candidate cells needs (drb_extract_workspace_bytes)."""
    return int(_lib.load().drb_extract_workspace_bytes(int(n_occupied)))


@torch.no_grad()
def extract_block(radiance_field, sample_grid, occupancy_binary, meta_data, device, jitter=None, workspace=None):
    """Evaluator.sample_points (eval_ngp_nerf.py:337-412) without the .ply side products:
    -> (voxel_grid float32 [R,R,R,7], voxel_mask int64 [K]) ready for torch.save."""
    sample_grid.set_binary_fields(occupancy_binary)
    pts, rgb, alpha, indices, dmask, smask, grid = sample_grid.query_radiance_and_density_from_camera(
        radiance_field, occupancy_binary, meta_data, device, jitter=jitter, return_grid=True,
        surface_only_where_dense=True, rgb_only_where_masked=True, workspace=workspace)
    return grid, indices[dmask & smask]
