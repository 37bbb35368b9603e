"""Import shim that lets the reference's own pure-PyTorch modules run on CPU.

TEST INFRASTRUCTURE.  Only usable where /root/reference exists (the build
container); never on the GPU box.  It inserts empty stand-ins for the CUDA-only
third-party wheels the reference imports at module scope (MinkowskiEngine,
torch_scatter, nerfacc, tinycudann, trimesh) and swaps the MinkowskiEngine based
``hierarchical_grid_subsample`` (conerf/register/grid_downsample.py:47-94) for
the oracle's deterministic restatement, because ME cannot be installed here.
Everything else (FPN, transformer, decoder, Procrustes) is the reference's code.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DREG_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "MinkowskiEngine", "torch_scatter", "nerfacc", "nerfacc.cuda", "nerfacc.contraction",
    "nerfacc.grid", "nerfacc.intersection", "nerfacc.vol_rendering", "nerfacc.pack",
    "tinycudann", "trimesh",
]


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "conerf"))


def import_reference():
    """Returns the reference's ``conerf.register.nerf_regtr`` module (CPU usable)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in _STUBS:
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = []  # behave like a package for "from a.b import c"
            sys.modules[name] = mod
    # names pulled by "from x import y" statements at module scope
    sys.modules["torch_scatter"].scatter_max = None
    sys.modules["nerfacc"].rendering = None
    sys.modules["nerfacc"].OccupancyGrid = None
    sys.modules["nerfacc"].ContractionType = None
    sys.modules["nerfacc.contraction"].ContractionType = None
    sys.modules["nerfacc.contraction"].contract_inv = None
    sys.modules["nerfacc.grid"].Grid = None
    sys.modules["nerfacc.intersection"].ray_aabb_intersect = None
    sys.modules["nerfacc.vol_rendering"]._RenderingTransmittanceFromAlphaCUB = None
    sys.modules["nerfacc.vol_rendering"]._RenderingTransmittanceFromAlphaNaive = None
    sys.modules["nerfacc.pack"].pack_info = None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    nerf_regtr = importlib.import_module("conerf.register.nerf_regtr")
    from oracle.downsample import hierarchical_grid_subsample as _hgs
    nerf_regtr.hierarchical_grid_subsample = _hgs
    return nerf_regtr
