"""dreg-nerf_b200: B200-native registration hot path of DReg-NeRF behind the reference's API.

Only what the path needs: ``csrc/`` (hand-written sm_100a CUDA + the C ABI of
``include/dregb200.h``), the ctypes binding, and the host-side mirrors of the reference
interfaces (``NeRFRegTr``, ``NGPradianceField`` / ``SampleGrid`` extract).
"""
from ._lib import DrbError, load as load_library  # noqa: F401
from .nerf_regtr import NeRFRegTr  # noqa: F401
from .ngp import NGPradianceField, SampleGrid, extract_block, extract_workspace_bytes  # noqa: F401
from . import augment, blockio, losses, pipeline, synthetic  # noqa: F401
from .pipeline import PairPipeline  # noqa: F401
from .occupancy import OccupancyGrid  # noqa: F401
from .optim import FusedAdamW  # noqa: F401
from .losses import CorrespondenceLoss, InfoNCELoss, RegistrationLoss  # noqa: F401
from .confidence_loss import compute_visibility_score, surface_field_mask  # noqa: F401
