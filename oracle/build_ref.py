"""Recipe for oracle/_ref/: the reference's OWN implementation of the hot path, byte-compiled.

comprna/riser is pure Python, so "compiling the reference from its sources where they lie" is
``py_compile``: the modules of the read-classification path under /root/reference/riser are compiled to
sourceless byte-code files (``.rbc`` = a .pyc under another suffix: gpurun's snapshot, like most sync tools, leaves
``*.pyc`` behind) in oracle/_ref/ (git-ignored, NOT gpurun-ignored: like a built .so it travels to the
GPU box, where /root/reference does not exist).  No reference source text enters the repository or its history;
oracle/_ref/MANIFEST.json records the SHA-256 of every source file that was compiled.

Used by (and only by) test infrastructure: ``oracle/refshim.py`` imports it when /root/reference is absent, so that
``bench.py --impl reference`` / ``cpu_baseline`` time the unmodified reference (np.vectorize normalise,
riser/preprocess.py:108-147; Model.classify, riser/model.py:22-28) on the GPU box's host cores, and the
"PyTorch on B200" bar runs the reference's own ConvNet module.

    python -m oracle.build_ref            # no-op when /root/reference is absent (the GPU box uses the prebuilt files)
"""
import hashlib
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("RISER_REFERENCE", "/root/reference")

# module path under riser/ -> path under oracle/_ref/ (the reference uses flat imports: `from nets.cnn import ...`)
MODULES = ["preprocess.py", "model.py", "control.py", "nets/cnn.py", "nets/resnet.py", "retrain/preprocess.py"]
SUFFIX = ".rbc"
CONFIGS = ["model/mRNA_config_RNA002_R9.4.1.yaml"]      # hyper-parameters only (riser/model/*.yaml:6-12)


def build(verbose=False):
    src_root = os.path.join(REF_ROOT, "riser")
    if not os.path.isdir(src_root):
        return None
    manifest = {"python": sys.version.split()[0], "magic": __import__("importlib.util").util.MAGIC_NUMBER.hex(),
                "files": {}}
    for rel in MODULES:
        src = os.path.join(src_root, rel)
        dst = os.path.join(OUT, rel[:-3] + SUFFIX)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: what tracebacks show instead of a path that does not exist on the GPU box
        py_compile.compile(src, cfile=dst, dfile=f"<comprna/riser>/riser/{rel}", doraise=True, optimize=0)
        with open(src, "rb") as f:
            manifest["files"][rel] = hashlib.sha256(f.read()).hexdigest()
        if verbose:
            print("compiled", rel)
    # the shipped hyper-parameters as JSON (data, not code): what get_config() would read
    import yaml
    for rel in CONFIGS:
        with open(os.path.join(src_root, rel)) as f:
            cfg = yaml.safe_load(f)
        with open(os.path.join(OUT, os.path.basename(rel).replace(".yaml", ".json")), "w") as f:
            json.dump(cfg, f)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return OUT


if __name__ == "__main__":
    print(build(verbose=True) or f"{REF_ROOT} absent: nothing to do")
