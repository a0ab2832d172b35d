"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference (Peer222/art-sbir) from
/root/reference so its own functions can pin the oracle and mint golden vectors.

Only usable in the build container (the GPU box has no /root/reference); callers must check
`available()` first.  Four third-party modules the reference imports but never uses for
arithmetic (torchinfo, matplotlib, seaborn, bresenham) are absent from the image and are
stubbed; importing `visualization` creates ./visual/, so the import runs
from a temporary directory.  Nothing under art_sbir_b200/ imports this file.
"""
from __future__ import annotations

import contextlib
import os
import sys
import tempfile
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("SBIR_REFERENCE_ROOT", "/root/reference"))
_cache = {}


def available() -> bool:
    return (REFERENCE_ROOT / "inference.py").is_file() and (REFERENCE_ROOT / "utils.py").is_file()


class _Anything(types.ModuleType):
    """Module stub: any attribute is a no-op callable / sub-stub."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(f"{self.__name__}.{name}")
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return self


def _install_stubs():
    for name in ("torchinfo", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.image",
                 "matplotlib.ticker", "matplotlib.colors", "seaborn", "bresenham", "py7zr"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Anything(name)


@contextlib.contextmanager
def _in_tempdir():
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            yield
        finally:
            os.chdir(old)


def load():
    """Returns (utils, inference) modules of the reference, imported once."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    _install_stubs()
    saved = {k: sys.modules.get(k) for k in ("utils", "inference", "models", "visualization", "data_preparation")}
    sys.path.insert(0, str(REFERENCE_ROOT))
    try:
        with _in_tempdir():
            import inference as ref_inference  # noqa: E402  (reference module)
            import utils as ref_utils  # noqa: E402
    finally:
        sys.path.remove(str(REFERENCE_ROOT))
    # keep the reference's modules out of the way of same-named modules elsewhere
    for k, v in saved.items():
        if v is not None:
            sys.modules[k] = v
    _cache["mods"] = (ref_utils, ref_inference)
    return _cache["mods"]
