"""Recipe for ``oracle/_ref``: the UNMODIFIED reference implementation of the path, staged so it can travel.

The reference (steb6/ISBFSAR) is pure Python -- there is nothing to compile.  This recipe copies, byte for byte,
the few files that define the scoring path from where they lie under ``/root/reference`` into ``oracle/_ref/``
(same relative layout, so the reference's own ``from modules.ar.utils.model import TRXOS`` /
``from utils.params import TRXConfig`` imports resolve with ``oracle/_ref`` on ``sys.path``):

    modules/ar/utils/model.py      TRXOS and everything it calls          (the hot path)
    modules/ar/ar.py               ActionRecognizer                        (stateful wrapper)
    utils/params.py                TRXConfig                               (configuration)
    modules/hpe/utils/misc.py      reconstruct_absolute / is_within_fov    (decode helpers)
    assets/saved/support_set.pkl   the reference's saved support set       (data fixture, main.py:321-333)

``oracle/_ref/`` is git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so the copy
travels to the GPU box, where ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg time the
reference's own ``TRXOS.forward`` on the host cores (``cpu_baseline.kind = "reference"``).  Without ``oracle/_ref``
those legs fall back to the oracle port (``kind = "port"``).  A SHA-256 manifest is written next to the copies and
checked at import time by ``oracle/ref_runner.py``.

    python -m oracle.build_ref            # run in the build container (needs /root/reference)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("ARX_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = [
    "modules/ar/utils/model.py",
    "modules/ar/ar.py",
    "utils/params.py",
    "modules/hpe/utils/misc.py",
    "assets/saved/support_set.pkl",
]


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def build(force: bool = False) -> str | None:
    """Stage the reference files; returns the output directory, or None when the reference tree is absent
    (e.g. on the GPU box, which only uses an already-staged copy)."""
    if not os.path.isdir(REF):
        return OUT if os.path.exists(os.path.join(OUT, "MANIFEST.json")) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or sha256(dst) != sha256(src):
            shutil.copyfile(src, dst)
        manifest[rel] = sha256(dst)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": "steb6/ISBFSAR (unmodified files, staged by oracle/build_ref.py)", "sha256": manifest}, f, indent=1)
    return OUT


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out if out else "reference tree not found at " + REF)
