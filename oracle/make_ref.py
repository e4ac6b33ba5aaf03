"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference for machines without /root/reference.

    python oracle/make_ref.py

The reference path (``/root/reference/mg/model/MusicTransformer``) is pure Python, so "building" it means
byte-compiling its own source files where they lie: every module the path imports is compiled with
``py_compile`` straight from ``/root/reference`` into ``oracle/_ref/MusicTransformer/<name>.bytecode`` (pyc format; the extension keeps snapshot tools that drop ``*.pyc`` from losing it).  No
reference SOURCE is copied into this repository; ``oracle/_ref/`` is a build output (git-ignored, but it
travels to the GPU box with the snapshot like the built ``.so``).  ``oracle/ref_import.py`` loads the
compiled modules when the source tree is absent, which is what lets ``bench.py --impl reference`` and the
``cpu_baseline`` / ``gpu_eager_baseline`` legs time the reference's own code (``kind: "reference"``) on the
B200 host instead of the oracle port.

``MANIFEST.json`` records the interpreter magic and the SHA-256 of every source file the bytecode came from.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import py_compile
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("MT_REFERENCE_DIR") or "/root/reference/mg/model/MusicTransformer"
OUT = os.path.join(ROOT, "oracle", "_ref", "MusicTransformer")
# import order of the reference's bare-name modules (MT/network.py:1-11, MT/utils.py:1-7, MT/data.py:1-8)
MODULES = ["sequence", "utils", "config", "layers", "criterion", "parallel", "metrics", "network", "data"]


def build(verbose: bool = True) -> str | None:
    """Returns the staging directory, or None when the reference sources are not present here."""
    if not os.path.isfile(os.path.join(SRC, "layers.py")):
        return None
    os.makedirs(OUT, exist_ok=True)
    manifest = {"source_dir": SRC, "python": sys.version.split()[0],
                "magic": importlib.util.MAGIC_NUMBER.hex(), "modules": {}}
    for name in MODULES:
        src = os.path.join(SRC, name + ".py")
        with open(src, "rb") as f:
            digest = hashlib.sha256(f.read()).hexdigest()
        # dfile: tracebacks keep pointing at the reference file the bytecode was compiled from
        py_compile.compile(src, cfile=os.path.join(OUT, name + ".bytecode"), dfile=src, doraise=True)
        manifest["modules"][name] = digest
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if verbose:
        print(f"staged {len(MODULES)} compiled reference modules in {OUT}")
    return OUT


if __name__ == "__main__":
    if build() is None:
        print(f"reference sources not found under {SRC}; nothing staged")
        sys.exit(1)
