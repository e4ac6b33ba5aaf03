"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference MusicTransformer modules.

Only usable where ``/root/reference`` exists (the build container).  Nothing under
``musicgeneration_b200/`` may import this file; only ``tests/``, ``oracle/make_golden.py``
do (the reference tree does not exist on the GPU box).

The reference (``mg/model/MusicTransformer``) is a flat script directory that imports its
siblings by bare name (``import utils``, ``import config`` -- network.py:1-2,7) and pulls in
three packages that are absent from this image and are not on the numeric path
(``pretty_midi`` via sequence.py:5, ``tensorboardX`` and ``progress.bar`` via network.py:10-11).
We register inert stubs for those three, then import the reference files under private module
names so they cannot clash with our own ``utils``/``config``.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

_CANDIDATES = [
    os.environ.get("MT_REFERENCE_DIR", ""),
    "/root/reference/mg/model/MusicTransformer",
]
# byte-compiled copy of the same files (oracle/make_ref.py): what exists on the GPU box
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "MusicTransformer")


def reference_dir() -> str | None:
    """Directory of the reference SOURCES (build container only)."""
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "layers.py")):
            return os.path.abspath(c)
    return None


def staged_dir() -> str | None:
    """Directory of the compiled reference modules staged by oracle/make_ref.py, if built for this interpreter."""
    man = os.path.join(_STAGED, "MANIFEST.json")
    if not os.path.isfile(man):
        return None
    try:
        import json
        with open(man) as f:
            magic = json.load(f).get("magic")
    except Exception:
        return None
    return _STAGED if magic == importlib.util.MAGIC_NUMBER.hex() else None


def reference_available() -> bool:
    return reference_dir() is not None or staged_dir() is not None


def _install_stubs() -> None:
    if "pretty_midi" not in sys.modules:
        m = types.ModuleType("pretty_midi")
        for n in ("PrettyMIDI", "Instrument", "ControlChange"):
            setattr(m, n, type(n, (), {"__init__": lambda self, *a, **k: None}))

        class Note:                 # the four fields MT/sequence.py reads and writes (pretty_midi.Note's signature)
            def __init__(self, velocity, pitch, start, end):
                self.velocity, self.pitch, self.start, self.end = velocity, pitch, start, end

        m.Note = Note
        sys.modules["pretty_midi"] = m
    if "tensorboardX" not in sys.modules:
        m = types.ModuleType("tensorboardX")
        m.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda self, *a, **k: None})
        sys.modules["tensorboardX"] = m
    if "progress" not in sys.modules:
        p = types.ModuleType("progress")
        b = types.ModuleType("progress.bar")

        class Bar:  # progress.bar.Bar(...).iter(it)
            def __init__(self, *a, **k):
                pass

            def iter(self, it):
                return it

        b.Bar = Bar
        p.bar = b
        sys.modules["progress"] = p
        sys.modules["progress.bar"] = b


_CACHE: dict | None = None


def load_reference() -> types.SimpleNamespace:
    """Return a namespace with the reference modules: layers, network, criterion, utils,
    config, metrics, data.  Raises FileNotFoundError when the reference tree is absent."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    d = reference_dir()
    ext = ".py"
    if d is None:
        d, ext = staged_dir(), ".bytecode"
    if d is None:
        raise FileNotFoundError("reference MusicTransformer not found (neither the sources nor the compiled "
                                "modules of oracle/make_ref.py)")
    _install_stubs()
    # The reference's bare-name imports need its directory on sys.path while loading; we
    # import under the bare names (that is what the files themselves do) but snapshot and
    # restore any modules of ours that share those names.
    bare = ["sequence", "utils", "config", "layers", "criterion", "parallel", "metrics",
            "network", "data"]
    saved = {n: sys.modules.pop(n) for n in bare if n in sys.modules}
    sys.path.insert(0, d)
    try:
        mods = {}
        for n in bare:
            path = os.path.join(d, n + ext)
            loader = importlib.machinery.SourcelessFileLoader(n, path) if ext != ".py" else None
            spec = importlib.util.spec_from_file_location(n, path, loader=loader)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[n] = mod
            spec.loader.exec_module(mod)
            mods[n] = mod
    finally:
        sys.path.remove(d)
        for n in bare:
            m = sys.modules.pop(n, None)
            if m is not None:
                sys.modules["_mtref_" + n] = m
        sys.modules.update(saved)
    # the reference modules look each other up through their own globals (already bound),
    # so removing the bare names from sys.modules is safe.
    _CACHE = types.SimpleNamespace(dir=d, compiled=(ext != ".py"), **mods)
    return _CACHE
