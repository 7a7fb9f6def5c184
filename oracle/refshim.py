"""Import the real reference (comprna/riser) read-only from /root/reference.

In the build container the sources under /root/reference are imported; on the GPU box (no
/root/reference) the byte-compiled copy oracle/build_ref.py left in oracle/_ref/ is imported instead.
Used by tests/golden/make_golden.py to produce the committed fixtures, by the tests that put the reference
next to riser_b200, and by bench.py's CPU / "PyTorch on B200" baseline legs.  Test infrastructure only.
Recipe: SURVEY.md appendix A.3 -- stub the three dead imports, put riser/ on
sys.path (the reference uses flat imports, riser/model.py:3)."""
import os
import sys
import types

REF_ROOT = os.environ.get("RISER_REFERENCE", "/root/reference")
BUILT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")    # oracle/build_ref.py's output


def source_available():
    return os.path.isdir(os.path.join(REF_ROOT, "riser"))


def built_available():
    return os.path.exists(os.path.join(BUILT, "preprocess.rbc"))


def _load_built(name, rel):
    """Import one byte-compiled reference module (oracle/_ref/<rel>.rbc) under the flat name the reference's own
    import statements use (`import preprocess`, `from nets.cnn import ConvNet`)."""
    import importlib.machinery
    import importlib.util
    if name in sys.modules:
        return sys.modules[name]
    loader = importlib.machinery.SourcelessFileLoader(name, os.path.join(BUILT, rel + ".rbc"))
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def available():
    """The reference can be imported: from its sources (build container) or from the byte-compiled copy that
    oracle/build_ref.py leaves in oracle/_ref/ (travels to the GPU box)."""
    return source_available() or built_available()


def kind():
    return "source" if source_available() else ("byte-compiled" if built_available() else None)


class AttrDict(dict):
    """Stand-in for ``attridict`` (riser/riser.py:21-23)."""
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) else v


def _stub_dead_imports():
    for name in ("matplotlib", "matplotlib.pyplot", "torchinfo"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["torchinfo"].summary = lambda *a, **k: None


def load():
    """-> namespace with preprocess, model, control, ConvNet, ResNet modules/classes."""
    if not available():
        raise RuntimeError(f"reference present neither at {REF_ROOT} nor byte-compiled in {BUILT}")
    _stub_dead_imports()
    if source_available():
        path = os.path.join(REF_ROOT, "riser")
        if path not in sys.path:
            sys.path.insert(0, path)
        import preprocess, model, control          # noqa: E401
        from nets.cnn import ConvNet
        from nets.resnet import ResNet
    else:
        if "nets" not in sys.modules:
            pkg = types.ModuleType("nets")
            pkg.__path__ = []
            sys.modules["nets"] = pkg
        ConvNet = _load_built("nets.cnn", "nets/cnn").ConvNet
        ResNet = _load_built("nets.resnet", "nets/resnet").ResNet
        preprocess = _load_built("preprocess", "preprocess")
        model = _load_built("model", "model")
        control = _load_built("control", "control")
    return types.SimpleNamespace(preprocess=preprocess, model=model, control=control,
                                 ConvNet=ConvNet, ResNet=ResNet)


def load_retrain_preprocess():
    """riser/retrain/preprocess.py (its ont_fast5_api import is stubbed: only the arithmetic
    functions are used)."""
    import importlib.machinery
    import importlib.util
    for name in ("ont_fast5_api", "ont_fast5_api.fast5_interface"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["ont_fast5_api.fast5_interface"].get_fast5_file = lambda *a, **k: None
    if source_available():
        spec = importlib.util.spec_from_file_location(
            "riser_retrain_preprocess", os.path.join(REF_ROOT, "riser", "retrain", "preprocess.py"))
    else:
        path = os.path.join(BUILT, "retrain", "preprocess.rbc")
        spec = importlib.util.spec_from_loader(
            "riser_retrain_preprocess", importlib.machinery.SourcelessFileLoader("riser_retrain_preprocess", path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cnn_config():
    """riser/model/*_config_*.yaml:6-12 (identical in all shipped configs)."""
    if source_available():
        import yaml
        with open(os.path.join(REF_ROOT, "riser", "model", "mRNA_config_RNA002_R9.4.1.yaml")) as f:
            return AttrDict(yaml.safe_load(f))
    import json
    with open(os.path.join(BUILT, "mRNA_config_RNA002_R9.4.1.json")) as f:
        return AttrDict(json.load(f))
