"""Copies the reference's nanoGPT scripts BYTE FOR BYTE into baseline/_ref/nanoGPT/ (git-ignored, not gpurun-ignored: it
travels to the GPU box with the snapshot but never enters the history).  Run in the build container, where /root/reference
exists; `__graft_entry__.build()` calls it.  Used by
  * tools/bench_torch_ref.py — the same-box PyTorch GPU number (the reference GPT, eager and torch.compile);
  * bench.py --impl reference / cpu_baseline — the unmodified reference step on the host cores (kind "reference").
Nothing in the product package imports it."""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/nanoGPT"
DST = os.path.join(ROOT, "baseline", "_ref", "nanoGPT")
FILES = ["model.py", "configurator.py", "bench.py", "train.py", "sample.py"]


def vendor() -> bool:
    if not os.path.isdir(SRC):
        return os.path.exists(os.path.join(DST, "model.py"))
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        assert filecmp.cmp(os.path.join(SRC, f), os.path.join(DST, f), shallow=False)
    return True


def import_reference_model():
    """(GPT, GPTConfig) of the vendored, unmodified reference model.py, or None if baseline/_ref is absent."""
    path = os.path.join(DST, "model.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_nanogpt_model", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.GPT, mod.GPTConfig


if __name__ == "__main__":
    ok = vendor()
    print("baseline/_ref/nanoGPT:", "ready" if ok else "unavailable (no /root/reference and no earlier copy)")
    sys.exit(0 if ok else 1)
