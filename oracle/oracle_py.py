"""ctypes view of oracle/liboracle.so (the CPU restatement).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"
REF_BIN = HERE / "_ref" / "shafa"

NSYM = 256
MAXC = 256


class OrcBuf(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("size", C.c_uint64), ("cap", C.c_uint64)]


class OrcFResult(C.Structure):
    _fields_ = [
        ("rle_mode", C.c_int),
        ("n_blocks", C.c_uint64),
        ("rle", OrcBuf),
        ("rle_freq", OrcBuf),
        ("freq", OrcBuf),
        ("rle_sizes", C.POINTER(C.c_uint64)),
    ]


def build(force: bool = False) -> None:
    """Compile the restatement (gcc only).  Building the checker is not using it."""
    src = HERE / "shafa_oracle.c"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(HERE), "all"], stdout=subprocess.DEVNULL)
    if not REF_BIN.exists() and Path("/root/reference/src").is_dir():
        subprocess.check_call(["make", "-C", str(HERE), "_ref"], stdout=subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
        L.orc_rle_encode_block.restype = C.c_uint64
        L.orc_rle_encode_block.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_rle_gate.restype = C.c_int
        L.orc_rle_gate.argtypes = [C.c_uint64, C.c_uint64, C.c_int]
        L.orc_histogram.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_freq_fields.restype = C.c_size_t
        L.orc_freq_fields.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_freq_parse.argtypes = [C.c_char_p, C.c_void_p]
        L.orc_sf_codes.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_sf_codes_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_cod_fields.restype = C.c_size_t
        L.orc_cod_fields.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_cod_parse.argtypes = [C.c_char_p, C.c_void_p]
        L.orc_sf_encode_block.restype = C.c_uint64
        L.orc_sf_encode_block.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_sf_encoded_size.restype = C.c_uint64
        L.orc_sf_encoded_size.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_sf_decode_block.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_rle_decode_block.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, u64p]
        L.orc_module_f.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(OrcFResult)]
        L.orc_f_result_free.argtypes = [C.POINTER(OrcFResult)]
        L.orc_module_t.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(OrcBuf)]
        L.orc_module_c.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(OrcBuf)]
        L.orc_module_d_shaf.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(OrcBuf)]
        L.orc_module_d_rle.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(OrcBuf)]
        L.orc_buf_free.argtypes = [C.POINTER(OrcBuf)]
        _ = u8p
        _lib = L
    return _lib


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(a, dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _take(buf: OrcBuf) -> bytes:
    out = C.string_at(buf.data, buf.size) if buf.size else b""
    lib().orc_buf_free(C.byref(buf))
    return out


class OracleError(RuntimeError):
    def __init__(self, code: int):
        super().__init__(f"oracle error {code}")
        self.code = code


# ---------------- per-block ----------------

def rle_encode_block(block) -> np.ndarray:
    a = _u8(block)
    out = np.empty(2 * a.size + 3, dtype=np.uint8)
    n = lib().orc_rle_encode_block(_ptr(a), a.size, _ptr(out))
    return out[:n].copy()


def rle_gate(n0: int, rle0: int, force_rle: bool = False) -> bool:
    return bool(lib().orc_rle_gate(n0, rle0, int(force_rle)))


def histogram(block) -> np.ndarray:
    a = _u8(block)
    h = np.zeros(NSYM, dtype=np.uint64)
    lib().orc_histogram(_ptr(a), a.size, _ptr(h))
    return h


def freq_fields(hist) -> bytes:
    h = np.ascontiguousarray(hist, dtype=np.uint64)
    out = C.create_string_buffer(NSYM * 22)
    n = lib().orc_freq_fields(_ptr(h), out)
    return out.raw[:n]


def freq_parse(text: bytes) -> np.ndarray:
    h = np.zeros(NSYM, dtype=np.uint64)
    rc = lib().orc_freq_parse(text, _ptr(h))
    if rc:
        raise OracleError(rc)
    return h


def sf_codes(hist) -> list[str]:
    h = np.ascontiguousarray(hist, dtype=np.uint64)
    codes = np.zeros((NSYM, MAXC), dtype=np.uint8)
    lib().orc_sf_codes(_ptr(h), _ptr(codes))
    return [bytes(row[: int(np.argmax(row == 0))]).decode() if row[0] else "" for row in codes]


def sf_codes_packed(hist):
    h = np.ascontiguousarray(hist, dtype=np.uint64)
    ln = np.zeros(NSYM, dtype=np.uint8)
    cd = np.zeros(NSYM, dtype=np.uint64)
    mx = lib().orc_sf_codes_packed(_ptr(h), _ptr(ln), _ptr(cd))
    return ln, cd, mx


def _codes_array(codes: list[str]) -> np.ndarray:
    arr = np.zeros((NSYM, MAXC), dtype=np.uint8)
    for s, c in enumerate(codes):
        if c:
            arr[s, : len(c)] = np.frombuffer(c.encode(), dtype=np.uint8)
    return arr


def cod_fields(codes: list[str]) -> bytes:
    return ";".join(codes).encode()


def sf_encode_block(block, codes: list[str]) -> np.ndarray:
    a = _u8(block)
    arr = _codes_array(codes)
    h = histogram(a)
    z = lib().orc_sf_encoded_size(_ptr(h), _ptr(arr))
    out = np.zeros(z + 8, dtype=np.uint8)
    z2 = lib().orc_sf_encode_block(_ptr(a), a.size, _ptr(arr), _ptr(out))
    assert z2 == z
    return out[:z].copy()


def sf_decode_block(payload, n: int, codes: list[str]) -> np.ndarray:
    p = _u8(payload)
    arr = _codes_array(codes)
    out = np.empty(n, dtype=np.uint8)
    rc = lib().orc_sf_decode_block(_ptr(p), p.size, n, _ptr(arr), _ptr(out))
    if rc:
        raise OracleError(rc)
    return out


def rle_decode_block(block, cap: int | None = None) -> np.ndarray:
    a = _u8(block)
    fin = C.c_uint64(0)
    rc = lib().orc_rle_decode_block(_ptr(a), a.size, None, 0, C.byref(fin))
    if rc:
        raise OracleError(rc)
    out = np.empty(fin.value, dtype=np.uint8)
    rc = lib().orc_rle_decode_block(_ptr(a), a.size, _ptr(out), out.size, C.byref(fin))
    if rc:
        raise OracleError(rc)
    return out


# ---------------- whole file ----------------

def module_f(data, block_size: int, force_rle: bool = False, force_freq: bool = False) -> dict:
    a = _u8(data)
    res = OrcFResult()
    rc = lib().orc_module_f(_ptr(a), a.size, block_size, int(force_rle), int(force_freq), C.byref(res))
    if rc:
        raise OracleError(rc)
    B = res.n_blocks
    out = {
        "rle_mode": bool(res.rle_mode),
        "n_blocks": int(B),
        "rle": C.string_at(res.rle.data, res.rle.size) if res.rle.size else b"",
        "rle_freq": C.string_at(res.rle_freq.data, res.rle_freq.size) if res.rle_freq.size else b"",
        "freq": C.string_at(res.freq.data, res.freq.size) if res.freq.size else b"",
        "rle_sizes": [int(res.rle_sizes[i]) for i in range(B)] if res.rle_mode else [],
    }
    lib().orc_f_result_free(C.byref(res))
    return out


def module_t(freq_text: bytes) -> bytes:
    buf = OrcBuf()
    rc = lib().orc_module_t(freq_text, len(freq_text), C.byref(buf))
    if rc:
        raise OracleError(rc)
    return _take(buf)


def module_c(sym, cod_text: bytes) -> bytes:
    a = _u8(sym)
    buf = OrcBuf()
    rc = lib().orc_module_c(_ptr(a), a.size, cod_text, len(cod_text), C.byref(buf))
    if rc:
        raise OracleError(rc)
    return _take(buf)


def module_d_shaf(shaf: bytes, cod_text: bytes, rle: bool) -> bytes:
    buf = OrcBuf()
    rc = lib().orc_module_d_shaf(shaf, len(shaf), cod_text, len(cod_text), int(rle), C.byref(buf))
    if rc:
        raise OracleError(rc)
    return _take(buf)


def module_d_rle(rle: bytes, freq_text: bytes) -> bytes:
    buf = OrcBuf()
    rc = lib().orc_module_d_rle(rle, len(rle), freq_text, len(freq_text), C.byref(buf))
    if rc:
        raise OracleError(rc)
    return _take(buf)


# ---------------- the unmodified reference binary ----------------

def ref_available() -> bool:
    return REF_BIN.exists() and os.access(REF_BIN, os.X_OK)


def ref_run(args: list[str], cwd: str, timeout: float = 600.0) -> subprocess.CompletedProcess:
    """Run oracle/_ref/shafa (the reference's own main) with CLI args."""
    return subprocess.run([str(REF_BIN)] + args, cwd=cwd, capture_output=True, timeout=timeout)
