"""The `shafa` host binary: option handling on the CPU, and (GPU) file-level parity
with the reference's artefacts, including cross-decoding of the reference's files."""
from __future__ import annotations

import json
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, encode_only_cases, small_cases

BIN = ROOT / "shafa-cd_b200" / "shafa"


def run(args, cwd):
    return subprocess.run([str(BIN)] + args, cwd=cwd, capture_output=True, timeout=600)


def test_binary_built():
    assert BIN.exists(), "run `python __graft_entry__.py` (build()) first"


def test_option_errors(tmp_path):
    (tmp_path / "x").write_bytes(b"a" * 2000)
    for args, msg in [([], b"No file input"), (["-m", "f"], b"No file input"), (["x", "-m", "q"], b"Wrong Options' syntax"),
                      (["x", "-b", "Z"], b"Wrong Options' syntax"), (["x", "y"], b"Wrong Options' syntax"),
                      (["x", "-m"], b"Wrong Options' syntax"), (["x", "-m", "ff"], b"Wrong Options' syntax")]:
        r = run(args, tmp_path)
        assert r.returncode == 1 and msg in r.stderr, (args, r.stderr)
    # extension / ordering checks happen before any module runs (shafa.c:165-257)
    for args, msg in [(["x", "-m", "t"], b"Module t: Wrong extension"), (["x", "-m", "d", "-d", "s"], b"Module d: Wrong extension"),
                      (["x", "-m", "d", "-d", "r"], b"Module d: Wrong extension"), (["x", "-m", "d"], b"Module d: Wrong extension")]:
        r = run(args, tmp_path)
        assert r.returncode == 1 and msg in r.stderr, (args, r.stderr)


def test_missing_file_and_small_file(tmp_path):
    r = run(["nope", "-m", "f"], tmp_path)
    assert r.returncode == 1 and b"File can't be accessed" in r.stderr
    (tmp_path / "tiny").write_bytes(b"a" * 100)
    r = run(["tiny", "-m", "f"], tmp_path)
    assert r.returncode == 1 and b"File too small" in r.stderr   # f.c:220,366


def _has_gpu(pkg):
    return pkg.load_library().shafa_cuda_device_count() > 0


def test_no_gpu_is_an_error_not_a_fallback(pkg, tmp_path):
    if _has_gpu(pkg):
        pytest.skip("a GPU is present")
    (tmp_path / "x").write_bytes(bytes(range(256)) * 8)
    r = run(["x", "-m", "f"], tmp_path)
    assert r.returncode == 1 and not (tmp_path / "x.freq").exists()


@pytest.mark.gpu
@pytest.mark.parametrize("case", small_cases())
def test_cli_matches_reference_files(case, tmp_path):
    d = GOLDEN / case
    flags = json.loads((d / "flags.json").read_text())
    original = (d / "input.bin").read_bytes()
    (tmp_path / "input.bin").write_bytes(original)
    assert run(["input.bin", "-m", "f"] + flags["fflags"], tmp_path).returncode == 0
    rle = (tmp_path / "input.bin.rle").exists()
    base = "input.bin.rle" if rle else "input.bin"
    assert run([base + ".freq", "-m", "t"], tmp_path).returncode == 0
    assert run([base, "-m", "c"], tmp_path).returncode == 0
    if rle and (tmp_path / "input.bin.freq").exists():
        assert run(["input.bin.freq", "-m", "t"], tmp_path).returncode == 0
    produced = sorted(p.name for p in tmp_path.iterdir())
    expected = sorted(p.name for p in d.iterdir() if p.name != "flags.json")
    assert produced == expected
    for name in expected:
        assert (tmp_path / name).read_bytes() == (d / name).read_bytes(), name
    # decoding (the decoder overwrites its targets, like the reference)
    if rle:
        (tmp_path / "input.bin.rle").unlink()
        assert run(["input.bin.rle.shaf", "-m", "d", "-d", "s"], tmp_path).returncode == 0
        assert (tmp_path / "input.bin.rle").read_bytes() == (d / "input.bin.rle").read_bytes()
        (tmp_path / "input.bin").unlink()
        assert run(["input.bin.rle", "-m", "d", "-d", "r"], tmp_path).returncode == 0
        assert (tmp_path / "input.bin").read_bytes() == original
        (tmp_path / "input.bin").unlink()
        assert run(["input.bin.rle.shaf"], tmp_path).returncode == 0          # default: -m d on *.shaf (shafa.c:293-295)
        assert (tmp_path / "input.bin").read_bytes() == original
    else:
        (tmp_path / "input.bin").unlink()
        assert run(["input.bin.shaf", "-m", "d"], tmp_path).returncode == 0
        assert (tmp_path / "input.bin").read_bytes() == original


@pytest.mark.gpu
def test_cli_default_pipeline_and_block_flags(pkg, orc, tmp_path):
    """`shafa file -b K` = F,T,C in one invocation (shafa.c:293-298); compared with the oracle."""
    data = pkg.datagen.mixed4m(9 * 1024 * 1024 + 333, seed=5)
    (tmp_path / "f").write_bytes(data.tobytes())
    assert run(["f", "-b", "K"], tmp_path).returncode == 0
    f = orc.module_f(data, 655360)
    assert f["rle_mode"] and (tmp_path / "f.rle").read_bytes() == f["rle"]
    assert (tmp_path / "f.rle.freq").read_bytes() == f["rle_freq"]
    cod = orc.module_t(f["rle_freq"])
    assert (tmp_path / "f.rle.cod").read_bytes() == cod
    assert (tmp_path / "f.rle.shaf").read_bytes() == orc.module_c(f["rle"], cod)
    (tmp_path / "f").unlink()
    assert run(["f.rle.shaf", "-m", "d"], tmp_path).returncode == 0
    assert (tmp_path / "f").read_bytes() == data.tobytes()


@pytest.mark.gpu
def test_cross_decode_with_reference_binary(pkg, orc, tmp_path):
    """Our encoder -> reference decoder and reference encoder -> our decoder (where oracle/_ref exists)."""
    if not orc.ref_available():
        pytest.skip("oracle/_ref/shafa not present on this box")
    data = pkg.datagen.textlike(1_500_000, seed=77)
    ours, theirs = tmp_path / "ours", tmp_path / "theirs"
    ours.mkdir(); theirs.mkdir()
    (ours / "x").write_bytes(data.tobytes()); (theirs / "x").write_bytes(data.tobytes())
    assert run(["x", "-b", "K"], ours).returncode == 0
    for step in (["x", "-m", "f", "-b", "K"], ["x.rle.freq", "-m", "t"], ["x.rle", "-m", "c"]):
        assert orc.ref_run(step, str(theirs)).returncode == 0
    for name in ("x.rle", "x.rle.freq", "x.rle.cod", "x.rle.shaf"):
        assert (ours / name).read_bytes() == (theirs / name).read_bytes(), name
    (ours / "x").unlink(); (theirs / "x").unlink()
    assert orc.ref_run(["x.rle.shaf", "-m", "d"], str(ours)).returncode == 0     # reference decodes our files
    assert run(["x.rle.shaf", "-m", "d"], theirs).returncode == 0                 # we decode the reference's files
    assert (ours / "x").read_bytes() == data.tobytes() and (theirs / "x").read_bytes() == data.tobytes()
    _ = shutil


@pytest.mark.gpu
@pytest.mark.parametrize("case", encode_only_cases())
def test_cli_one_symbol_block(case, tmp_path):
    """`shafa x` on a file whose last block holds one distinct symbol exits 0 and writes the reference's bytes."""
    d = GOLDEN / "encode_only" / case
    flags = json.loads((d / "flags.json").read_text())
    (tmp_path / "input.bin").write_bytes((d / "input.bin").read_bytes())
    assert run(["input.bin"] + flags["fflags"], tmp_path).returncode == 0      # F, T, C in one invocation
    expected = sorted(p.name for p in d.iterdir() if p.name != "flags.json")
    assert sorted(p.name for p in tmp_path.iterdir()) == expected
    for name in expected:
        assert (tmp_path / name).read_bytes() == (d / name).read_bytes(), name
    shaf = [n for n in expected if n.endswith(".shaf")][0]
    assert run([shaf, "-m", "d"], tmp_path).returncode == 1                     # refused, not a crash


REFDRIVER = ROOT / "oracle" / "_ref" / "shafa_refdriver"


def test_reference_driver_links_against_our_modules():
    """the reference's unmodified shafa.c + extensions.c + errors.c link against
    host/mod_*.c + util.c + ref_shim.c + libshafa_cuda (CPU-only check: it links and parses options)."""
    if not Path("/root/reference/src/shafa.c").exists():
        pytest.skip("reference sources not on this box (the binary is prebuilt in the build container)")
    subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "refdriver"], stdout=subprocess.DEVNULL)
    assert REFDRIVER.exists()
    r = subprocess.run([str(REFDRIVER)], capture_output=True, timeout=60)
    assert r.returncode == 1 and b"No file input" in r.stderr                    # shafa.c:269-272
    r = subprocess.run([str(REFDRIVER), "x", "-m", "q"], capture_output=True, timeout=60)
    assert r.returncode == 1 and b"Wrong Options" in r.stderr                    # shafa.c:277-280


@pytest.mark.gpu
@pytest.mark.parametrize("case", small_cases())
def test_reference_driver_binding_matches_reference_files(case, tmp_path):
    """The golden cases through the REFERENCE's driver bound to our modules."""
    if not REFDRIVER.exists():
        pytest.skip("oracle/_ref/shafa_refdriver not present on this box")
    d = GOLDEN / case
    flags = json.loads((d / "flags.json").read_text())
    original = (d / "input.bin").read_bytes()
    (tmp_path / "input.bin").write_bytes(original)

    def ref(args):
        return subprocess.run([str(REFDRIVER)] + args, cwd=tmp_path, capture_output=True, timeout=600)

    assert ref(["input.bin", "-m", "f", "-m", "t", "-m", "c"] + flags["fflags"]).returncode == 0
    rle = (tmp_path / "input.bin.rle").exists()
    if rle and (tmp_path / "input.bin.freq").exists():
        assert ref(["input.bin.freq", "-m", "t"]).returncode == 0
    expected = sorted(p.name for p in d.iterdir() if p.name != "flags.json")
    assert sorted(p.name for p in tmp_path.iterdir()) == expected
    for name in expected:
        assert (tmp_path / name).read_bytes() == (d / name).read_bytes(), name
    (tmp_path / "input.bin").unlink()
    shaf = "input.bin.rle.shaf" if rle else "input.bin.shaf"
    assert ref([shaf, "--no-multithread"]).returncode == 0                       # default -m d; the flag is accepted
    assert (tmp_path / "input.bin").read_bytes() == original


@pytest.mark.gpu
def test_cli_fused_equals_separate_modules_and_multi_device(pkg, tmp_path):
    """`shafa f` (F, T, C fused in one pass) writes the same files as the three modules run one by one
    (SHAFA_NO_FUSE=1) and as the same invocation sharded over several contexts (SHAFA_DEVICES)."""
    import os
    data = pkg.datagen.mixed4m(21 * 1024 * 1024 + 999, seed=31)
    outs = {}
    for tag, env in (("fused", {}), ("separate", {"SHAFA_NO_FUSE": "1"}), ("multi", {"SHAFA_DEVICES": "0,0,0"})):
        d = tmp_path / tag
        d.mkdir()
        (d / "f").write_bytes(data.tobytes())
        r = subprocess.run([str(BIN), "f", "-b", "K", "-c", "f"], cwd=d, capture_output=True, timeout=600, env={**os.environ, **env})
        assert r.returncode == 0, r.stderr
        outs[tag] = {p.name: p.read_bytes() for p in d.iterdir() if p.name != "f"}
        assert sorted(outs[tag]) == ["f.freq", "f.rle", "f.rle.cod", "f.rle.freq", "f.rle.shaf"]
        (d / "f").unlink()
        r = subprocess.run([str(BIN), "f.rle.shaf", "-m", "d"], cwd=d, capture_output=True, timeout=600, env={**os.environ, **env})
        assert r.returncode == 0 and (d / "f").read_bytes() == data.tobytes()
        (d / "f").unlink()
        r = subprocess.run([str(BIN), "f.rle", "-m", "d", "-d", "r"], cwd=d, capture_output=True, timeout=600, env={**os.environ, **env})
        assert r.returncode == 0 and (d / "f").read_bytes() == data.tobytes()
    assert outs["fused"] == outs["separate"] == outs["multi"]


@pytest.mark.gpu
def test_cli_codes_longer_than_64_bits(pkg, orc, tmp_path):
    """`-m c` and `-m d` with a hand-made .cod whose codes reach 255 bits (c.c:18 keeps 32 bytes per code, d.c:422-504 builds
    its tree from the strings): byte-identical to the unmodified reference's .shaf (tests/golden/wide_codes), and the R-mode
    route (wide SF decode, then RLE expansion) against the oracle."""
    d = GOLDEN / "wide_codes"
    for name in ("x", "x.cod"):
        shutil.copy(d / name, tmp_path / name)
    r = run(["x", "-m", "c"], tmp_path)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "x.shaf").read_bytes() == (d / "x.shaf").read_bytes()
    (tmp_path / "x").unlink()
    r = run(["x.shaf", "-m", "d"], tmp_path)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "x").read_bytes() == (d / "x").read_bytes()
    # R mode: our own F, then a hand-made wide table over the symbols of the .rle stream
    data = pkg.datagen.runheavy(300_000, seed=4)
    (tmp_path / "y").write_bytes(data.tobytes())
    r = run(["y", "-m", "f", "-c", "r"], tmp_path)
    assert r.returncode == 0, r.stderr
    rle = (tmp_path / "y.rle").read_bytes()
    freq = (tmp_path / "y.rle.freq").read_bytes()
    sizes = [int(x) for x in freq.split(b"@")[3:-1:2]]
    present = sorted(set(rle))
    assert len(present) <= 255
    codes = [""] * 256
    for k, sym in enumerate(present[:-1]):
        codes[sym] = "1" * (k * (254 // max(1, len(present) - 1))) + "0" if k else "0"
    depth_last = max(len(c) for c in codes)
    codes[present[-1]] = "1" * depth_last
    assert 64 < depth_last <= 255
    cod = f"@R@{len(sizes)}".encode() + b"".join(f"@{sz}@".encode() + ";".join(codes).encode() for sz in sizes) + b"@0"
    (tmp_path / "y.rle.cod").write_bytes(cod)
    r = run(["y.rle", "-m", "c"], tmp_path)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "y.rle.shaf").read_bytes() == orc.module_c(rle, cod)
    (tmp_path / "y").unlink()
    (tmp_path / "y.rle").unlink()
    r = run(["y.rle.shaf", "-m", "d"], tmp_path)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "y").read_bytes() == data.tobytes()
