"""Differential test: oracle restatement vs the unmodified reference binary
(oracle/_ref/shafa) on fresh random inputs.  Skipped where the binary is absent."""
from __future__ import annotations

import numpy as np
import pytest


def _ref_pipeline(orc, tmp_path, data, bflag, fflags):
    (tmp_path / "x").write_bytes(data.tobytes())
    b = ["-b", bflag] if bflag else []
    assert orc.ref_run(["x", "-m", "f"] + b + fflags, str(tmp_path)).returncode == 0
    rle = (tmp_path / "x.rle").exists()
    base = "x.rle" if rle else "x"
    assert orc.ref_run([base + ".freq", "-m", "t"], str(tmp_path)).returncode == 0
    assert orc.ref_run([base, "-m", "c"], str(tmp_path)).returncode == 0
    return rle, base


@pytest.mark.parametrize("seed", range(6))
def test_random_against_reference(orc, pkg, tmp_path, seed):
    if not orc.ref_available():
        pytest.skip("oracle/_ref/shafa not built here")
    rng = np.random.default_rng(1000 + seed)
    gen = ["zipf", "textlike", "runheavy", "uniform", "mixed4m", "runheavy"][seed]
    n = int(rng.integers(1024, 400_000))
    data = pkg.datagen.GENERATORS[gen](n, seed=seed)
    if seed == 5:  # sprinkle zeros and long zero runs
        data = data.copy()
        data[rng.integers(0, n, size=n // 50)] = 0
        data[n // 2: n // 2 + 700] = 0
    fflags = [[], [], ["-c", "f"], ["-c", "r"], [], ["-c", "r"]][seed]
    rle, base = _ref_pipeline(orc, tmp_path, data, None, fflags)
    f = orc.module_f(data, 65536, force_rle="r" in fflags[1::2], force_freq="f" in fflags[1::2])
    assert f["rle_mode"] == rle
    sym = f["rle"] if rle else data.tobytes()
    freq = f["rle_freq"] if rle else f["freq"]
    assert (tmp_path / (base + ".freq")).read_bytes() == freq
    if rle:
        assert (tmp_path / "x.rle").read_bytes() == f["rle"]
    cod = orc.module_t(freq)
    assert (tmp_path / (base + ".cod")).read_bytes() == cod
    assert (tmp_path / (base + ".shaf")).read_bytes() == orc.module_c(sym, cod)
