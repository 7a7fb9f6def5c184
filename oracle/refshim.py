"""Import the real reference (comprna/riser) read-only from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
tests/golden/make_golden.py to produce the committed fixtures and by
tests/test_oracle_vs_reference.py (skipped when the reference is absent).
Recipe: SURVEY.md appendix A.3 -- stub the three dead imports, put riser/ on
sys.path (the reference uses flat imports, riser/model.py:3)."""
import os
import sys
import types

REF_ROOT = os.environ.get("RISER_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "riser"))


class AttrDict(dict):
    """Stand-in for ``attridict`` (riser/riser.py:21-23)."""
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) else v


def load():
    """-> namespace with preprocess, model, control, ConvNet, ResNet modules/classes."""
    if not available():
        raise RuntimeError("reference not present at " + REF_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "torchinfo"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["torchinfo"].summary = lambda *a, **k: None
    path = os.path.join(REF_ROOT, "riser")
    if path not in sys.path:
        sys.path.insert(0, path)
    import preprocess, model, control          # noqa: E401
    from nets.cnn import ConvNet
    from nets.resnet import ResNet
    return types.SimpleNamespace(preprocess=preprocess, model=model, control=control,
                                 ConvNet=ConvNet, ResNet=ResNet)


def load_retrain_preprocess():
    """riser/retrain/preprocess.py (its ont_fast5_api import is stubbed: only the arithmetic
    functions are used)."""
    import importlib.util
    for name in ("ont_fast5_api", "ont_fast5_api.fast5_interface"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["ont_fast5_api.fast5_interface"].get_fast5_file = lambda *a, **k: None
    spec = importlib.util.spec_from_file_location(
        "riser_retrain_preprocess", os.path.join(REF_ROOT, "riser", "retrain", "preprocess.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cnn_config():
    """riser/model/*_config_*.yaml:6-12 (identical in all shipped configs)."""
    import yaml
    with open(os.path.join(REF_ROOT, "riser", "model", "mRNA_config_RNA002_R9.4.1.yaml")) as f:
        return AttrDict(yaml.safe_load(f))
