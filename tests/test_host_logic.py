"""CPU-only tests: the C-ABI library loads and exports every symbol include/arx.h declares, the
reference-shaped host classes keep the state_dict schema, and the sharding helpers are right."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle.synth import Cfg, make_state_dict
from tests.util import Args, torch_sd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from isbfsar_b200 import _lib
    lib_path = _lib.LIB_PATH
    if not os.path.exists(lib_path):
        from isbfsar_b200.build import build
        build()
    hdr = open(os.path.join(ROOT, "include", "arx.h")).read()
    declared = sorted(set(re.findall(r"ARX_API[^;(]*?\b(arx_\w+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in arx.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert _lib.load().arx_abi_version() == _lib.ABI_VERSION


def test_state_dict_schema_matches_reference():
    from isbfsar_b200 import TRXOS, TRXConfig
    m = TRXOS(TRXConfig())
    expect = {
        "features_extractor.sk.fc1.weight": (180, 90), "features_extractor.sk.fc1.bias": (180,),
        "features_extractor.sk.fc2.weight": (256, 180), "features_extractor.sk.fc2.bias": (256,),
        "transformers.0.pe.pe": (1, 24, 256),
        "transformers.0.k_linear.weight": (128, 512), "transformers.0.k_linear.bias": (128,),
        "transformers.0.v_linear.weight": (128, 512), "transformers.0.v_linear.bias": (128,),
        "transformers.0.norm_k.weight": (128,), "transformers.0.norm_k.bias": (128,),
        "discriminator.dimensionality_reduction.weight": (16, 128), "discriminator.dimensionality_reduction.bias": (16,),
        "discriminator.fc1.weight": (256, 1920), "discriminator.fc1.bias": (256,),
        "discriminator.fc2.weight": (64, 256), "discriminator.fc2.bias": (64,),
        "discriminator.fc3.weight": (1, 64), "discriminator.fc3.bias": (1,),
        "post_resnet.l1.weight": (256, 2048), "post_resnet.l1.bias": (256,),
    }   # SURVEY.md 8b, probed from the reference
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == expect
    sd = make_state_dict(Cfg(), 0, include_post_resnet=True)
    m.load_state_dict(torch_sd(sd))                                   # strict
    assert np.array_equal(m.transformers[0].pe.pe.numpy(), sd["transformers.0.pe.pe"])
    assert m.transformers[0].tuples_len == 120 and len(m.transformers[0].tuples) == 120
    assert m.transformers[0].scores == []


def test_no_cpu_fallback():
    from isbfsar_b200 import TRXOS
    m = TRXOS(Args(Cfg()))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m({"sk": torch.zeros(1, 5, 16, 90)}, torch.arange(5)[None], {"sk": torch.zeros(1, 16, 90)})


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "isbfsar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src, f


@pytest.mark.parametrize("n,world", [(0, 2), (1, 2), (7, 2), (8, 4), (65536, 8), (10, 3)])
def test_shard_bounds_partition(n, world):
    from isbfsar_b200.dist import shard_bounds
    prev = 0
    for r in range(world):
        s, e = shard_bounds(n, world, r)
        assert s == prev and e >= s
        prev = e
    assert prev == n
    sizes = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
