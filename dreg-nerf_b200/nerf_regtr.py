"""Drop-in ``NeRFRegTr`` backed by libdregb200 (hand-written sm_100a CUDA).

Mirrors the reference's module surface (conerf/register/nerf_regtr.py:72-248): same constructor
arguments, same ``forward(data) -> dict`` contract, same 772-entry ``state_dict`` (including the
``fpn3d.feature_pyramid.resnet.*`` alias of ``fpn3d.backbone_net.*``,
conerf/model/feature_pyramid_net.py:43,194-200), so checkpoints written by the reference's
``CheckPointManager`` load unchanged.  The sub-modules below only *hold* parameters in the
reference's layout; all arithmetic happens in the C-ABI engine.  There is no PyTorch fallback.
"""
import copy
import ctypes as C
import os
import threading

import torch
from torch import nn

from . import _lib

_PRECISION_PLANES = {"fp32": 2, "bf16x3": 2, "bf16": 1}
# attention through the tcgen05 FlashAttention-style kernel (csrc/attention.cu) instead of the mma.sync one
TC_ATTENTION_DEFAULT = os.environ.get("DRB_TC_ATTENTION", "1") != "0"


# Engine slot of the calling thread.  An engine's workspaces belong to the work queued on ONE stream, so pairs that
# are in flight on different CUDA streams (pipeline.PairPipeline: one worker thread + stream per slot) each need an
# engine of their own; slot 0 is the ordinary single-stream use.
_tls = threading.local()
_ENGINE_LOCK = threading.Lock()     # engine creation (module level: the module stays deep-copyable / picklable)


def engine_slot():
    return getattr(_tls, "slot", 0)


def set_engine_slot(slot):
    _tls.slot = int(slot)


# ------------------------------------------------------------------------------------------------
# parameter holders (construction order follows the reference so that torch.manual_seed(s) followed
# by NeRFRegTr() yields the reference's initial weights)
# ------------------------------------------------------------------------------------------------
class _Bottleneck(nn.Module):
    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv3d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm3d(planes)
        self.conv2 = nn.Conv3d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm3d(planes)
        self.conv3 = nn.Conv3d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm3d(planes * 4)
        self.downsample = downsample


class _ResNet3D50(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.conv1 = nn.Conv3d(in_channels, 64, kernel_size=5, stride=2, padding=2, bias=False)
        self.bn1 = nn.BatchNorm3d(64)
        inplanes = 64
        for li, (planes, nblk) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3))):
            stride = 1 if li == 0 else 2
            down = nn.Sequential(nn.Conv3d(inplanes, planes * 4, kernel_size=1, stride=stride, bias=False),
                                 nn.BatchNorm3d(planes * 4))
            blocks = [_Bottleneck(inplanes, planes, stride, down)]
            inplanes = planes * 4
            blocks += [_Bottleneck(inplanes, planes, 1, None) for _ in range(1, nblk)]
            setattr(self, "layer%d" % (li + 1), nn.Sequential(*blocks))
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.xavier_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()


def _fpn_conv(cin, cout, k):
    layer = nn.Conv3d(cin, cout, kernel_size=k, padding=k // 2)
    nn.init.xavier_normal_(layer.weight)
    nn.init.constant_(layer.bias.data, val=0)
    return layer


class _FeaturePyramid(nn.Module):
    def __init__(self, resnet):
        super().__init__()
        self.resnet = resnet
        self.pyramid_transformation_1 = _fpn_conv(64, 256, 3)
        self.pyramid_transformation_2 = _fpn_conv(256, 256, 1)
        self.pyramid_transformation_3 = _fpn_conv(512, 256, 1)
        self.pyramid_transformation_4 = _fpn_conv(1024, 256, 1)
        self.pyramid_transformation_5 = _fpn_conv(2048, 256, 1)
        for i in range(1, 5):
            setattr(self, "upsample_transform_%d" % i, _fpn_conv(256, 256, 3))


class _FPN3D(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.backbone_net = _ResNet3D50(in_channels)
        self.feature_pyramid = _FeaturePyramid(self.backbone_net)


class _EncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_ff):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=0.0)
        self.cross_attn = nn.MultiheadAttention(d_model, nhead, dropout=0.0)
        self.linear1 = nn.Linear(d_model, dim_ff)
        self.linear2 = nn.Linear(dim_ff, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)


class _CrossEncoder(nn.Module):
    def __init__(self, layer, num_layers, norm):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])
        self.norm = norm


class _SinePosEmbed(nn.Module):
    def __init__(self, scale):
        super().__init__()
        self.scale = scale


class _Decoder(nn.Module):
    def __init__(self, d_model, pos_embed):
        super().__init__()
        self.pos_embed = pos_embed
        self.q_norm = nn.LayerNorm(d_model)     # declared but unused by the reference forward
        self.q_proj = nn.Linear(d_model, d_model)
        self.k_proj = nn.Linear(d_model, d_model)
        self.conf_logits_decoder = nn.Linear(d_model, 1)


# ------------------------------------------------------------------------------------------------
class NeRFRegTr(nn.Module):
    """``NeRFRegTr(pos_emb_type='sine', pos_emb_dim=256, pos_emb_scaling=1.0, num_downsample=6)``.

    Extra keyword (not in the reference): ``precision`` - ``'fp32'`` (default; split-bf16 tensor-core
    products, fp32-grade results) or ``'bf16'`` (single bf16 product, the bf16 configs of
    BASELINE.json).
    """

    def __init__(self, pos_emb_type: str = "sine", pos_emb_dim: int = 256, pos_emb_scaling: float = 1.0,
                 num_downsample: int = 6, precision: str = "fp32") -> None:
        super().__init__()
        if pos_emb_type != "sine":
            raise NotImplementedError("libdregb200 implements the 'sine' positional embedding only "
                                      "(the reference's default, train_nerf_regtr.py:89-94)")
        if pos_emb_dim != 256:
            raise NotImplementedError("libdregb200 kernels are specialised for pos_emb_dim == 256")
        if precision not in _PRECISION_PLANES:
            raise ValueError("precision must be one of %s" % sorted(_PRECISION_PLANES))
        self.num_downsample = num_downsample
        self.pos_emb_scaling = float(pos_emb_scaling)
        self.precision = precision
        self.fpn3d = _FPN3D(in_channels=4)
        self.pos_embed = _SinePosEmbed(pos_emb_scaling)
        layer = _EncoderLayer(pos_emb_dim, 8, 1024)
        self.transformer_encoder = _CrossEncoder(layer, 6, nn.LayerNorm(pos_emb_dim))
        self.correspondence_decoder = _Decoder(pos_emb_dim, self.pos_embed)
        self._engines = {}
        # evaluate the two level-1 FPN convolutions only where the masked gather reads p1 (exact)
        self.sparse_fpn = True
        self.max_tokens = 3000
        self.last_token_counts = (0, 0)

    # -------------------------------------------------------------------------------------------
    def _named_tensors(self):
        """name -> tensor of every parameter and buffer (aliases included).  Walking the module tree costs ~2 ms of
        Python per call - a fifth of a 128^3 forward - so the dict is cached and dropped whenever the tensors may
        have been replaced (``_apply``: .to() / .cuda() / .float(); ``load_state_dict`` copies in place)."""
        cache = getattr(self, "_tensor_cache", None)
        if cache is None:
            cache = dict(self.named_parameters(remove_duplicate=False))
            cache.update(dict(self.named_buffers(remove_duplicate=False)))
            object.__setattr__(self, "_tensor_cache", cache)
        return cache

    def _apply(self, fn, *args, **kwargs):
        object.__setattr__(self, "_tensor_cache", None)
        out = super()._apply(fn, *args, **kwargs)
        object.__setattr__(self, "_tensor_cache", None)
        return out

    def invalidate_tensor_cache(self):
        """Call after replacing a Parameter / buffer object by assignment (rare; in-place updates need nothing)."""
        object.__setattr__(self, "_tensor_cache", None)

    def _get_engine(self, res_xyz, device, max_mask):
        slot = engine_slot()
        max_mask = max(int(max_mask), int(getattr(self, "min_mask_capacity", 0)))
        key = (tuple(res_xyz), device.index, self.precision, slot)
        ent = self._engines.get(key)
        if ent is not None and ent["max_mask"] >= max_mask:
            return ent
        with _ENGINE_LOCK:
            return self._create_engine(key, ent, res_xyz, device, max_mask, slot)

    def _create_engine(self, key, ent, res_xyz, device, max_mask, slot):
        lib = _lib.load()
        if ent is not None:
            lib.drb_engine_destroy(ent["handle"])
        cap = max(int(max_mask * 1.25) + 1024, 4096)
        cfg = _lib.EngineConfig(res_x=res_xyz[0], res_y=res_xyz[1], res_z=res_xyz[2],
                                planes=_PRECISION_PLANES[self.precision],
                                num_downsample=self.num_downsample,
                                pos_emb_scaling=self.pos_emb_scaling, max_mask=cap,
                                training_bn=1 if self.training else 0)
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.drb_engine_create(C.byref(cfg), C.byref(handle)), "drb_engine_create")
        names = [lib.drb_engine_param_name(handle, i).decode() for i in range(lib.drb_engine_num_params(handle))]
        # slot > 0: a secondary engine of a stream pipeline leaves the shared BatchNorm running buffers to slot 0
        _lib.check(lib.drb_engine_set_update_running(handle, 1 if slot == 0 else 0))
        ent = {"handle": handle, "names": names, "max_mask": cap, "sig": None, "device": device, "slot": slot}
        idx = [i for i in range(len(names)) if lib.drb_engine_param_trainable(handle, i)]
        numels = [int(lib.drb_engine_param_numel(handle, i)) for i in idx]
        offs, total = [], 0
        for n in numels:
            offs.append(total)
            total += (n + 3) // 4 * 4          # 16-byte aligned views (vector loads / stores in the kernels)
        ent.update(train_idx=idx, train_names=[names[i] for i in idx], train_numels=numels,
                   train_offsets=offs, train_total=total)
        if getattr(self, "_profile", False):
            _lib.check(lib.drb_engine_set_profile(handle, 1))
        self._engines[key] = ent
        return ent

    def _sync_params(self, ent):
        """(Re)binds and repacks the weights when any tensor was replaced or written to."""
        lib = _lib.load()
        tensors = self._named_tensors()
        bound = ent.get("bound_list")
        if bound is None or ent.get("bound_cache") is not tensors:
            bound = [tensors[n] for n in ent["names"]]
            ent["bound_list"], ent["bound_cache"] = bound, tensors
        sig = tuple((t.data_ptr(), t._version) for t in bound)
        if sig == ent["sig"]:
            return
        for i, n in enumerate(ent["names"]):
            t = tensors[n]
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                raise _lib.DrbError("parameter %s must be a contiguous fp32 CUDA tensor" % n)
            if t.numel() != lib.drb_engine_param_numel(ent["handle"], i):
                raise _lib.DrbError("parameter %s has %d elements, engine expects %d"
                                    % (n, t.numel(), lib.drb_engine_param_numel(ent["handle"], i)))
            _lib.check(lib.drb_engine_bind_param(ent["handle"], i, _lib.ptr(t)), "drb_engine_bind_param")
        _lib.check(lib.drb_engine_commit_params(ent["handle"], _lib.stream_ptr()), "drb_engine_commit_params")
        ent["sig"] = sig

    def launch_count(self):
        lib = _lib.load()
        return sum(int(lib.drb_engine_launch_count(e["handle"])) for e in self._engines.values())

    def set_profile(self, on):
        """Event-brackets every tensor-core GEMM launch (roofline instrumentation for bench.py)."""
        lib = _lib.load()
        self._profile = bool(on)
        for ent in self._engines.values():
            _lib.check(lib.drb_engine_set_profile(ent["handle"], int(on)))

    def read_profile(self):
        """-> (igemm device ms, algorithmic FLOPs, launches) since the last read."""
        lib = _lib.load()
        ms = fl = 0.0
        cnt = 0
        for ent in self._engines.values():
            a, b, c = C.c_double(0), C.c_double(0), C.c_longlong(0)
            _lib.check(lib.drb_engine_profile_read(ent["handle"], C.byref(a), C.byref(b), C.byref(c)))
            ms, fl, cnt = ms + a.value, fl + b.value, cnt + c.value
        return ms, fl, cnt

    def __del__(self):
        try:
            if not self._engines:          # a module that never ran owns nothing (and must not map the library)
                return
            lib = _lib.load()
            for ent in self._engines.values():
                lib.drb_engine_destroy(ent["handle"])
        except Exception:
            pass

    # -------------------------------------------------------------------------------------------
    def reserve_mask_capacity(self, n_masked):
        """Engines are (re)built when a pair brings more masked voxels than the engine was sized for (a few GB of
        workspace and a repack of the weights).  A caller that knows its largest pair - a data loader, a stream
        pipeline with one engine per slot - reserves it once and no engine is rebuilt mid-run."""
        self.min_mask_capacity = max(int(n_masked), int(getattr(self, "min_mask_capacity", 0)))

    def set_max_tokens(self, max_total):
        """Stopping rule of the down-sampler (grid_downsample.py:70,91 hard-code 3000 points for the pair);
        BASELINE.json's 8k-token configuration raises it."""
        self.max_tokens = int(max_total)

    def _pair_io(self, src, tgt, src_mask, tgt_mask):
        return _lib.PairIO(
            src_grid=src.data_ptr(), tgt_grid=tgt.data_ptr(),
            s_ch=src.stride(1), s_z=src.stride(2), s_x=src.stride(3), s_y=src.stride(4),
            t_ch=tgt.stride(1), t_z=tgt.stride(2), t_x=tgt.stride(3), t_y=tgt.stride(4),
            src_mask=src_mask.data_ptr(), n_src_mask=src_mask.numel(),
            tgt_mask=tgt_mask.data_ptr(), n_tgt_mask=tgt_mask.numel())

    @staticmethod
    def _pair_out(outs):
        src_feats, tgt_feats, src_kp, tgt_kp, src_corr, tgt_corr, src_ov, tgt_ov, pose = outs
        return _lib.PairOut(src_feats=src_feats.data_ptr(), tgt_feats=tgt_feats.data_ptr(),
                            src_kp=src_kp.data_ptr(), tgt_kp=tgt_kp.data_ptr(),
                            src_corr=src_corr.data_ptr(), tgt_corr=tgt_corr.data_ptr(),
                            src_overlap=src_ov.data_ptr(), tgt_overlap=tgt_ov.data_ptr(),
                            pose=pose.data_ptr())

    def _run_engine(self, ent, src, tgt, src_mask, tgt_mask, grad_mode):
        """encode + decode through the C ABI -> the nine output tensors."""
        lib = _lib.load()
        device = src.device
        with torch.cuda.device(device):
            handle = ent["handle"]
            _lib.check(lib.drb_engine_set_grad_mode(handle, 1 if grad_mode else 0))
            if grad_mode and not ent.get("grad_ready"):
                ent["sig"] = None            # the data-gradient weight planes are packed by the next commit
                ent["grad_ready"] = True
            self._sync_params(ent)
            _lib.check(lib.drb_engine_set_training(handle, 1 if self.training else 0))
            _lib.check(lib.drb_engine_set_sparse_fpn(handle, 1 if self.sparse_fpn else 0))
            _lib.check(lib.drb_engine_set_max_tokens(handle, int(getattr(self, "max_tokens", 3000))))
            _lib.check(lib.drb_engine_set_tc_attention(handle, 1 if getattr(self, "tc_attention", TC_ATTENTION_DEFAULT) else 0))
            io = self._pair_io(src, tgt, src_mask, tgt_mask)
            ns, nt = C.c_int(0), C.c_int(0)
            stream = _lib.stream_ptr()
            _lib.check(lib.drb_engine_encode(handle, C.byref(io), C.byref(ns), C.byref(nt), stream),
                       "drb_engine_encode")
            ns, nt = ns.value, nt.value
            self.last_token_counts = (ns, nt)
            f32 = dict(dtype=torch.float32, device=device)
            outs = (torch.empty((6, ns, 256), **f32), torch.empty((6, nt, 256), **f32),
                    torch.empty((ns, 3), **f32), torch.empty((nt, 3), **f32),
                    torch.empty((6, ns, 3), **f32), torch.empty((6, nt, 3), **f32),
                    torch.empty((6, ns, 1), **f32), torch.empty((6, nt, 1), **f32),
                    torch.empty((6, 1, 3, 4), **f32))
            out = self._pair_out(outs)
            _lib.check(lib.drb_engine_decode(handle, C.byref(out), stream), "drb_engine_decode")
            if self.training and ent.get("slot", 0) == 0:
                # nn.BatchNorm3d bookkeeping: two training-mode calls (src, tgt) per forward
                if getattr(self, "_nbt", None) is None or self._nbt[0].device != device:
                    self._nbt = [b for n, b in self.named_buffers() if n.endswith("num_batches_tracked")]
                torch._foreach_add_(self._nbt, 2)
        return outs

    def _run_backward(self, ent, inputs, outs, grads_out, need):
        """drb_engine_backward -> one gradient tensor (or None) per entry of ent['train_names']."""
        lib = _lib.load()
        src, tgt, src_mask, tgt_mask = inputs
        device = src.device
        with torch.cuda.device(device):
            handle = ent["handle"]
            numels = ent["train_numels"]
            offs = ent["train_offsets"]
            flat = torch.zeros(ent["train_total"], dtype=torch.float32, device=device)
            views = []
            for j, i in enumerate(ent["train_idx"]):
                v = flat[offs[j]:offs[j] + numels[j]]
                views.append(v)
                _lib.check(lib.drb_engine_bind_grad(handle, i, _lib.ptr(v) if need[j] else None))

            def g(t, shape):
                if t is None:
                    return None
                t = t.contiguous().float()
                assert tuple(t.shape) == tuple(shape), (t.shape, shape)
                return t
            d_sf, d_tf, _, _, d_sc, d_tc, d_so, d_to, d_pose = grads_out
            keep = [g(d_sf, outs[0].shape), g(d_tf, outs[1].shape), g(d_sc, outs[4].shape), g(d_tc, outs[5].shape),
                    g(d_so, outs[6].shape), g(d_to, outs[7].shape), g(d_pose, outs[8].shape)]
            pg = _lib.PairGrad(*[(t.data_ptr() if t is not None else None) for t in keep])
            io = self._pair_io(src, tgt, src_mask, tgt_mask)
            out = self._pair_out(outs)
            _lib.check(lib.drb_engine_backward(handle, C.byref(io), C.byref(out), C.byref(pg), _lib.stream_ptr()),
                       "drb_engine_backward")
            tensors = self._named_tensors()
            return [views[j].view(tensors[n].shape) if need[j] else None for j, n in enumerate(ent["train_names"])]

    def forward(self, data):
        """Same contract as conerf/register/nerf_regtr.py:112-248 (one pair per call).  When gradients are
        enabled and a parameter requires them, the outputs carry an autograd node whose backward runs the
        engine's backward pass (train_nerf_regtr.py:229)."""
        if len(data["src_xyz_rgba"].shape) == 6:
            data["src_xyz_rgba"] = data["src_xyz_rgba"].squeeze(0)
            data["tgt_xyz_rgba"] = data["tgt_xyz_rgba"].squeeze(0)
            data["src_mask"] = data["src_mask"].squeeze(0)
            data["tgt_mask"] = data["tgt_mask"].squeeze(0)
            if "src_nerf_path" in data:
                data["src_nerf_path"] = data["src_nerf_path"][0]
                data["tgt_nerf_path"] = data["tgt_nerf_path"][0]
            if "pose" in data:
                data["pose"] = data["pose"].squeeze(0)
        src, tgt = data["src_xyz_rgba"], data["tgt_xyz_rgba"]
        assert len(src.shape) == 5  # [batch_size, C, z_dim, x_dim, y_dim]
        if src.shape[0] != 1 or tgt.shape[0] != 1:
            raise _lib.DrbError("NeRFRegTr.forward processes one pair per call, as the reference does "
                                "(nerf_regtr.py:144-147 index batch element 0 only)")
        if not src.is_cuda:
            raise _lib.DrbError("libdregb200 has no CPU path: move the inputs to a CUDA device")
        if src.dtype != torch.float32 or tgt.dtype != torch.float32:
            raise _lib.DrbError("grids must be float32")
        if src.shape != tgt.shape or src.shape[1] != 7:
            raise _lib.DrbError("expected two [1, 7, Z, X, Y] grids of equal resolution")
        device = src.device
        src_mask = data["src_mask"].reshape(-1).to(device=device, dtype=torch.int64).contiguous()
        tgt_mask = data["tgt_mask"].reshape(-1).to(device=device, dtype=torch.int64).contiguous()
        _, _, Z, X, Y = src.shape
        with torch.cuda.device(device):
            ent = self._get_engine((X, Y, Z), device, max(src_mask.numel(), tgt_mask.numel()))
        tensors = self._named_tensors()
        if ent.get("train_cache") is not tensors:
            ent["train_params"], ent["train_cache"] = [tensors[n] for n in ent["train_names"]], tensors
        train_params = ent["train_params"]
        if torch.is_grad_enabled() and any(p.requires_grad for p in train_params):
            outs = _RegistrationFn.apply(self, ent, src, tgt, src_mask, tgt_mask, *train_params)
        else:
            outs = self._run_engine(ent, src, tgt, src_mask, tgt_mask, grad_mode=False)
        src_feats, tgt_feats, src_kp, tgt_kp, src_corr, tgt_corr, src_ov, tgt_ov, pose = outs
        return {
            "src_feats": [src_feats], "tgt_feats": [tgt_feats],
            "src_kp": [src_kp], "src_kp_warped": [src_corr],
            "tgt_kp": [tgt_kp], "tgt_kp_warped": [tgt_corr],
            "src_overlap": [src_ov], "tgt_overlap": [tgt_ov],
            "pose": pose,
        }

    def tap(self, name, which):
        """Debug / parity tap of an engine intermediate (channels-last fp32), see drb_engine_tap."""
        lib = _lib.load()
        ent = next(iter(self._engines.values()))
        cap = 1 << 28
        buf = torch.empty(cap, dtype=torch.float32, device=ent["device"])
        n = C.c_longlong(0)
        _lib.check(lib.drb_engine_tap(ent["handle"], name.encode(), which, _lib.ptr(buf), cap, C.byref(n),
                                      _lib.stream_ptr()), "drb_engine_tap")
        torch.cuda.synchronize()
        return buf[:n.value].clone()


class _RegistrationFn(torch.autograd.Function):
    """Autograd node of one NeRFRegTr.forward: forward = drb_engine_encode + decode with the training graph
    kept inside the engine, backward = drb_engine_backward (one backward per forward)."""

    @staticmethod
    def forward(ctx, module, ent, src, tgt, src_mask, tgt_mask, *params):
        outs = module._run_engine(ent, src, tgt, src_mask, tgt_mask, grad_mode=True)
        ctx.module, ctx.ent = module, ent
        ctx.need = [bool(p.requires_grad) for p in params]
        ctx.save_for_backward(src, tgt, src_mask, tgt_mask, *outs)
        ctx.mark_non_differentiable(outs[2], outs[3])       # key points depend on the inputs only
        return outs

    @staticmethod
    def backward(ctx, *grads_out):
        saved = ctx.saved_tensors
        inputs, outs = saved[:4], saved[4:]
        grads = ctx.module._run_backward(ctx.ent, inputs, outs, grads_out, ctx.need)
        return (None, None, None, None, None, None) + tuple(grads)
