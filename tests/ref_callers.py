"""Test helper: import the UNMODIFIED reference's own callers (train.forward, val_lm.visdial_evaluate) and model from a reference
checkout — ``baseline/_ref/reference`` (git-ignored copy made by scripts/make_ref_copy.py; travels to the GPU box) or
``/root/reference`` (build container).  Third-party modules the reference imports but that are absent offline are stubbed; none
of them does arithmetic on the hot path."""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(ROOT, "baseline", "_ref", "reference"), "/root/reference"]


def reference_root():
    for p in CANDIDATES:
        if os.path.exists(os.path.join(p, "train.py")):
            return p
    return None


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def import_reference(ref_root):
    """-> dict(train, val_lm, vd (models.vilbert_dialog), vde (models.visual_dialog_encoder), du (utils.data_utils))."""
    import torch
    _stub("pytorch_transformers")
    _stub("pytorch_transformers.modeling_bert", BertEmbeddings=object)
    _stub("pytorch_transformers.tokenization_bert", BertTokenizer=object)
    _stub("pytorch_transformers.optimization", AdamW=object)
    _stub("pytorch_pretrained_bert")
    _stub("pytorch_pretrained_bert.file_utils", cached_path=lambda *a, **k: None)
    for name in ("visdom", "h5py", "lmdb"):
        _stub(name)
    sys.modules["visdom"].Visdom = object
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self          # the reference calls .cuda() on a buffer it never reads
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    out = {}
    for key, mod in (("vd", "models.vilbert_dialog"), ("vde", "models.visual_dialog_encoder"), ("du", "utils.data_utils"),
                     ("vm", "utils.visdial_metrics"), ("train", "train"), ("val_lm", "val_lm")):
        out[key] = importlib.import_module(mod)
    return out


def build_reference_encoder(ref, ref_root, state_dict):
    """The reference's VisualDialogEncoder without its network download (as tests/golden/make_golden.py builds it)."""
    import torch
    vd, vde = ref["vd"], ref["vde"]
    cfg = vd.BertConfig.from_json_file(os.path.join(ref_root, "config", "bert_base_6layer_6conect.json"))
    model = vde.VisualDialogEncoder.__new__(vde.VisualDialogEncoder)
    torch.nn.Module.__init__(model)
    model.bert_pretrained = vd.BertForMultiModalPreTraining(cfg)
    model.load_state_dict(state_dict, strict=True)
    return model.eval()
