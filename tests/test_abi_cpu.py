"""CPU: the C-ABI library builds, loads, and exports every symbol include/rfmc_b200.h declares
(no compute call is made without a GPU); compute entry points refuse to run without a device."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rfmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rfmc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    from random_forest_mc_b200 import engine

    lib = engine.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert set(engine.EXPORTED_SYMBOLS) <= set(names)
    assert lib.rfmc_abi_version() == 3


def test_no_cpu_fallback_without_a_device():
    from random_forest_mc_b200 import engine

    if engine.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(engine.EngineUnavailable):
        engine.require_gpu()
    import numpy as np
    import pandas as pd

    from random_forest_mc_b200.model import RandomForestMC

    df = pd.DataFrame({"a": np.arange(40.0), "b": np.arange(40) % 3, "target": (np.arange(40) % 2).astype(str)})
    with pytest.raises(engine.EngineUnavailable):
        RandomForestMC(n_trees=2).fit(df)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "random-forest-mc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in src.splitlines():
                    code = line.split("//")[0].split("#")[0]
                    assert not re.search(r"\b(import|from)\s+oracle\b|oracle\.", code), (f, line)
