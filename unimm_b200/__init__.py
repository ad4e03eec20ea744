"""unimm_b200 — B200-native (sm_100a) generative-scoring hot path of UniMM-UL, behind the reference's
``VisualDialogEncoder`` boundary.  Importing the package loads ``lib/libunimm_b200.so``; there is no
CPU or PyTorch fallback (build it with ``__graft_entry__.build()``)."""
from .config import DEFAULT_CONFIG_PATH, ViLBertConfig, tiny_config  # noqa: F401
from .weights import param_shapes, random_state_dict  # noqa: F401


def __getattr__(name):
    # the CUDA-backed pieces are imported lazily so that config / weight utilities work without the .so
    if name in ("Engine", "HostArrays"):
        from . import engine
        return getattr(engine, name)
    if name == "VisualDialogEncoder":
        from .visual_dialog_encoder import VisualDialogEncoder
        return VisualDialogEncoder
    raise AttributeError(name)
