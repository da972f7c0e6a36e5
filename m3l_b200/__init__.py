"""m3l_b200 — B200-native (sm_100a) implementation of M3L's VTMAE/VTT train step.

    from m3l_b200 import VTT, VTMAE      # drop-in for models.pretrain_models.{VTT, VTMAE}
    from m3l_b200.vtt import VTT         # drop-in for models.VTT.VTT (DINO-side encoder), also m3l_b200.VTTDino
"""
__version__ = "0.1.0"

from ._lib import M3LError  # noqa: F401


def __getattr__(name):
    if name in ("VTT", "VTMAE", "EarlyCNN", "Transformer", "MAEExtractor", "pair"):
        from . import vtmae
        return getattr(vtmae, name)
    if name in ("DinoV2", "DinoCatMAEExtractor"):
        from . import dinov2
        return getattr(dinov2, name)
    if name in ("VTDINO", "DINOHead", "DINOLoss", "update_moving_average"):
        from . import vtdino
        return getattr(vtdino, name)
    if name == "VTTDino":
        from .vtt import VTT as VTTDino
        return VTTDino
    if name == "vt_load":
        from .data import vt_load
        return vt_load
    if name == "FusedTrainer":
        from .trainer import FusedTrainer
        return FusedTrainer
    raise AttributeError(name)
