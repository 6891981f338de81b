"""Import shim for the LIVE reference -- TEST INFRASTRUCTURE ONLY, build container only.

``/root/reference`` exists only in the build container (never on the GPU box), so
this module is used solely by ``tests/golden/make_golden.py`` and by the
``not gpu`` tests that re-check the oracle against the live code when it is
present.  Nothing is copied: the reference modules are imported from where they
lie, after

* stubbing ``matplotlib`` / ``ruptures`` (absent here; ``src/utils.py:1,4,9``
  import them at module top), and
* neutralising the hard-coded ``'cuda'`` (``src/model.py:36``,
  ``src/utils.py:90,119,137,141,143``, ``src/imported/labelprop.py:103``) so the
  code runs unmodified on CPU.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("CRW_REFERENCE_ROOT", "/root/reference")
# bench.py's reference arm times the reference on the HOST cores of a GPU box: the 'cuda' literals are mapped to CPU there too
FORCE_CPU = bool(os.environ.get("CRW_REFERENCE_FORCE_CPU"))


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "model.py"))


def _on_cpu() -> bool:
    return FORCE_CPU or not torch.cuda.is_available()


def _stub_modules():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            col = types.ModuleType("matplotlib.colors")
            col.ListedColormap = lambda *a, **k: None
            mpl.pyplot, mpl.colors = plt, col
            sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.colors": col})
    if "ruptures" not in sys.modules:
        try:
            import ruptures  # noqa: F401
        except Exception:
            sys.modules["ruptures"] = types.ModuleType("ruptures")


_loaded = {}


def load():
    """Returns a namespace with the reference modules: model, utils, labelprop, maskedatt, encoder, dataset."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"live reference not found under {REF_ROOT}")
    _stub_modules()
    src = os.path.join(REF_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    with open(os.devnull, "w") as devnull, contextlib.redirect_stdout(devnull):
        import model as ref_model
        import utils as ref_utils
        import encoder as ref_encoder
        import dataset as ref_dataset
        from imported import labelprop as ref_labelprop
        from imported import maskedatt as ref_maskedatt
    if _on_cpu():
        # model.py:36 -- zeros(..., device='cuda')
        ref_model.zeros = lambda *a, device=None, **k: torch.zeros(*a, **k)
    _loaded.update(model=ref_model, utils=ref_utils, encoder=ref_encoder, dataset=ref_dataset,
                   labelprop=ref_labelprop, maskedatt=ref_maskedatt)
    return types.SimpleNamespace(**_loaded)


@contextlib.contextmanager
def cpu_device_patches():
    """Map every hard-coded 'cuda' in ``propagate``/``predict`` to CPU for the duration."""
    if not _on_cpu():
        yield
        return
    orig_zeros, orig_to, orig_cuda = torch.zeros, torch.Tensor.to, torch.Tensor.cuda

    def zeros(*a, **k):
        if k.get("device") == "cuda":
            k["device"] = "cpu"
        return orig_zeros(*a, **k)

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x == "cuda") else x for x in a)
        if k.get("device") == "cuda":
            k["device"] = "cpu"
        return orig_to(self, *a, **k)

    torch.zeros = zeros
    torch.Tensor.to = to
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.zeros, torch.Tensor.to, torch.Tensor.cuda = orig_zeros, orig_to, orig_cuda


class TopkSpy:
    """Records (Ws, Is) of every ``batched_affinity`` call made by ``labelprop.predict``."""

    def __init__(self, ref):
        self.ref = ref
        self.W, self.I = [], []

    def __enter__(self):
        inner = self.ref.maskedatt.batched_affinity
        self._orig = self.ref.labelprop.batched_affinity

        def spy(*a, **k):
            Ws, Is = inner(*a, **k)
            self.W.append(Ws[0].detach().clone())
            self.I.append(Is[0].detach().clone())
            return Ws, Is

        self.ref.labelprop.batched_affinity = spy
        return self

    def __exit__(self, *exc):
        self.ref.labelprop.batched_affinity = self._orig
        return False
