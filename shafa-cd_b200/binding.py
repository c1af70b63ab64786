"""ctypes binding of libshafa_cuda (include/shafa_cuda.h) plus the text-container
helpers a host needs (.freq / .cod / .shaf framing, reference f.c:89-119,
t.c:350-361, c.c:256-258).

There is no CPU fallback: if the shared library is missing, lacks a declared
symbol, or no CUDA device is present, this raises.  The oracle is never imported
from here.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
import os as _os

LIB_PATH = Path(_os.environ["SHAFA_LIB"]) if _os.environ.get("SHAFA_LIB") else HERE / "libshafa_cuda.so"   # SHAFA_LIB: A/B builds of the same library
HEADER = HERE.parent / "include" / "shafa_cuda.h"

ERRORS = {
    0: "success", 1: "outside module", 2: "lack of memory", 3: "file inaccessible",
    4: "file unrecognizable", 5: "file stream failed", 6: "file too small",
    7: "thread creation failed / kernel launch failed", 8: "thread termination failed / sync failed",
}


class ShafaError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        super().__init__(f"libshafa_cuda error {code} ({ERRORS.get(code, '?')}) {detail}")
        self.code = code


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_uint64), ("total_ms", C.c_double)]


class EncodedC(C.Structure):
    _fields_ = [
        ("rle_mode", C.c_int), ("n_blocks", C.c_uint64), ("n", C.c_uint64), ("block_size", C.c_uint64),
        ("block_sizes", C.POINTER(C.c_uint64)),
        ("rle", C.POINTER(C.c_uint8)), ("rle_capacity", C.c_uint64), ("rle_total", C.c_uint64),
        ("sym_sizes", C.POINTER(C.c_uint64)),
        ("hist", C.POINTER(C.c_uint64)), ("hist_orig", C.POINTER(C.c_uint64)),
        ("code_len", C.POINTER(C.c_uint8)), ("code_bits", C.POINTER(C.c_uint64)),
        ("payload", C.POINTER(C.c_uint8)), ("payload_capacity", C.c_uint64), ("payload_total", C.c_uint64),
        ("payload_offsets", C.POINTER(C.c_uint64)), ("z", C.POINTER(C.c_uint64)),
        ("owns_rle", C.c_int), ("owns_payload", C.c_int),
    ]


def declared_symbols() -> list[str]:
    """Every function include/shafa_cuda.h declares."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(shafa_cuda_\w+)\s*\(", text)))


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ShafaError(7, f"{LIB_PATH} not built (run `python __graft_entry__.py`); there is no CPU fallback")
    L = C.CDLL(str(LIB_PATH))
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    if missing:
        raise ShafaError(7, f"library lacks declared symbols: {missing}")
    vp, u64, u64p, u8p = C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p
    L.shafa_cuda_device_count.restype = C.c_int
    L.shafa_cuda_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.shafa_cuda_destroy.argtypes = [vp]
    L.shafa_cuda_set_stream.argtypes = [vp, vp]
    L.shafa_cuda_synchronize.argtypes = [vp]
    L.shafa_cuda_last_error.argtypes = [vp]
    L.shafa_cuda_last_error.restype = C.c_char_p
    L.shafa_cuda_host_alloc.argtypes = [C.c_size_t]
    L.shafa_cuda_host_alloc.restype = vp
    L.shafa_cuda_host_free.argtypes = [vp]
    L.shafa_cuda_dev_alloc.argtypes = [vp, C.c_size_t]
    L.shafa_cuda_dev_alloc.restype = vp
    L.shafa_cuda_dev_free.argtypes = [vp, vp]
    L.shafa_cuda_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.shafa_cuda_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.shafa_cuda_profile_enable.argtypes = [vp, C.c_int]
    L.shafa_cuda_profile_reset.argtypes = [vp]
    L.shafa_cuda_profile_read.argtypes = [vp, C.POINTER(KernelStat), C.c_int, C.POINTER(C.c_int)]
    L.shafa_cuda_launch_count.argtypes = [vp]
    L.shafa_cuda_launch_count.restype = u64
    L.shafa_cuda_rle_gate.argtypes = [u64, u64, C.c_int]
    L.shafa_cuda_rle_encode_blocks.argtypes = [vp, u8p, u64, u64, u8p, u64, vp, vp]
    L.shafa_cuda_histogram_blocks.argtypes = [vp, u8p, vp, u64, vp]
    L.shafa_cuda_sf_build_codes.argtypes = [vp, vp, u64, vp, vp]
    L.shafa_cuda_sf_encode_blocks.argtypes = [vp, u8p, vp, u64, vp, vp, vp, u8p, u64, vp, vp]
    L.shafa_cuda_sf_decode_blocks.argtypes = [vp, u8p, vp, vp, vp, u64, vp, vp, u8p, u64]
    L.shafa_cuda_rle_decode_blocks.argtypes = [vp, u8p, vp, u64, u8p, u64, vp]
    L.shafa_cuda_encode_host.argtypes = [vp, vp, u64, u64, C.c_int, C.c_int, C.POINTER(EncodedC)]
    L.shafa_cuda_encode_shard_host.argtypes = [vp, vp, u64, u64, C.c_int, C.c_int, C.POINTER(EncodedC)]
    L.shafa_cuda_encoded_free.argtypes = [C.POINTER(EncodedC)]
    L.shafa_cuda_decode_host.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp, C.c_int, vp, u64, u64p, vp]
    L.shafa_cuda_rle_decode_host.argtypes = [vp, vp, vp, u64, vp, u64, u64p, vp]
    L.shafa_cuda_sf_encode_wide_host.argtypes = [vp, vp, vp, u64, vp, vp, vp, u64, vp, vp, u64p]
    L.shafa_cuda_sf_decode_wide_host.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp, vp, u64]
    L.shafa_cuda_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.shafa_cuda_multi_destroy.argtypes = [vp]
    L.shafa_cuda_multi_device_count.argtypes = [vp]
    L.shafa_cuda_multi_ctx.argtypes = [vp, C.c_int]
    L.shafa_cuda_multi_ctx.restype = vp
    L.shafa_cuda_multi_last_error.argtypes = [vp]
    L.shafa_cuda_multi_last_error.restype = C.c_char_p
    L.shafa_cuda_multi_launch_count.argtypes = [vp]
    L.shafa_cuda_multi_launch_count.restype = u64
    L.shafa_cuda_encode_host_multi.argtypes = [vp, vp, u64, u64, C.c_int, C.c_int, C.POINTER(EncodedC)]
    L.shafa_cuda_decode_host_multi.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp, C.c_int, vp, u64, u64p, vp]
    L.shafa_cuda_rle_decode_host_multi.argtypes = [vp, vp, vp, u64, vp, u64, u64p, vp]
    _ = u64p
    _lib = L
    return L


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(a, dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


# ---------------------------------------------------------------------------
# text containers (host side; tiny)
# ---------------------------------------------------------------------------

def freq_fields(hist_row) -> bytes:
    """f.c:94-110: count printed when it differs from the previous symbol's."""
    out = []
    prev = None
    for v in hist_row.tolist():
        out.append("" if prev is not None and v == prev else str(v))
        prev = v
    return ";".join(out).encode()


def freq_text(mode: str, sizes, hist) -> bytes:
    """`.freq` / `.rle.freq` (f.c:289-294,312-314,112-116)."""
    B = len(sizes)
    parts = [f"@{mode}@{B}".encode()]
    for b in range(B):
        parts.append(f"@{int(sizes[b])}@".encode())
        parts.append(freq_fields(hist[b]))
    parts.append(b"@0")
    return b"".join(parts)


def code_strings(len_row, code_row) -> list[str]:
    return [format(int(c), "b").zfill(int(l)) if l else "" for l, c in zip(len_row.tolist(), code_row.tolist())]


def cod_text(mode: str, sizes, code_len, code_bits) -> bytes:
    """`.cod` (t.c:302,350-361,395-396)."""
    B = len(sizes)
    parts = [f"@{mode}@{B}".encode()]
    for b in range(B):
        parts.append(f"@{int(sizes[b])}@".encode())
        parts.append(";".join(code_strings(code_len[b], code_bits[b])).encode())
    parts.append(b"@0")
    return b"".join(parts)


def parse_cod(text: bytes):
    """`.cod` -> (mode, sizes[B], code_len[B,256] u8, code_bits[B,256] u64). Codes > 64 bits raise (error 4)."""
    m = re.match(rb"@(.)@(\d+)", text)
    if not m:
        raise ShafaError(4, "bad .cod header")
    mode, B = m.group(1).decode(), int(m.group(2))
    pos = m.end()
    sizes = np.zeros(B, dtype=np.uint64)
    ln = np.zeros((B, 256), dtype=np.uint8)
    cd = np.zeros((B, 256), dtype=np.uint64)
    for b in range(B):
        m = re.compile(rb"@(\d+)@([01;]*)").match(text, pos)
        if not m:
            raise ShafaError(5, ".cod block header")
        sizes[b] = int(m.group(1))
        fields = m.group(2).split(b";")
        if len(fields) != 256:
            raise ShafaError(4, ".cod block does not have 256 fields")
        for s, f in enumerate(fields):
            if len(f) > 64:
                raise ShafaError(4, "code longer than 64 bits")
            ln[b, s] = len(f)
            cd[b, s] = int(f, 2) if f else 0
        pos = m.end()
    return mode, sizes, ln, cd


def parse_cod_wide(text: bytes):
    """`.cod` with codes of up to 255 bits -> (mode, sizes[B], code_len[B,256] u8, code4[B,256,4] u64, word 0 least significant):
    the form of shafa_cuda_sf_encode_wide_host / shafa_cuda_sf_decode_wide_host."""
    m = re.match(rb"@(.)@(\d+)", text)
    if not m:
        raise ShafaError(4, "bad .cod header")
    mode, B = m.group(1).decode(), int(m.group(2))
    pos = m.end()
    sizes = np.zeros(B, dtype=np.uint64)
    ln = np.zeros((B, 256), dtype=np.uint8)
    c4 = np.zeros((B, 256, 4), dtype=np.uint64)
    for b in range(B):
        m = re.compile(rb"@(\d+)@([01;]*)").match(text, pos)
        if not m:
            raise ShafaError(5, ".cod block header")
        sizes[b] = int(m.group(1))
        fields = m.group(2).split(b";")
        if len(fields) != 256:
            raise ShafaError(4, ".cod block does not have 256 fields")
        for s, f in enumerate(fields):
            if len(f) > 255:
                raise ShafaError(4, "code longer than 255 bits")
            ln[b, s] = len(f)
            v = int(f, 2) if f else 0
            for k in range(4):
                c4[b, s, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
        pos = m.end()
    return mode, sizes, ln, c4


def shaf_bytes(z, payload: np.ndarray, payload_offsets) -> bytes:
    """`.shaf` (c.c:351,256-258): @B then @z@payload per block."""
    B = len(z)
    parts = [f"@{B}".encode()]
    for b in range(B):
        o, zz = int(payload_offsets[b]), int(z[b])
        parts.append(f"@{zz}@".encode())
        parts.append(payload[o:o + zz].tobytes())
    return b"".join(parts)


def parse_shaf(blob: bytes):
    """`.shaf` -> (z[B], payload (16-byte aligned slots), payload_offsets[B])."""
    m = re.match(rb"@(\d+)", blob)
    if not m:
        raise ShafaError(5, "bad .shaf header")
    B = int(m.group(1))
    pos = m.end()
    z = np.zeros(B, dtype=np.uint64)
    offs = np.zeros(B, dtype=np.uint64)
    chunks = []
    total = 0
    hdr = re.compile(rb"@(\d+)@")
    for b in range(B):
        m = hdr.match(blob, pos)
        if not m:
            raise ShafaError(5, ".shaf block header")
        zz = int(m.group(1))
        pos = m.end()
        if pos + zz > len(blob):
            raise ShafaError(5, ".shaf truncated")
        z[b] = zz
        offs[b] = total
        chunks.append((total, blob[pos:pos + zz]))
        pos += zz
        total += (zz + 15) & ~15
    payload = np.zeros(total + 16, dtype=np.uint8)
    for o, c in chunks:
        payload[o:o + len(c)] = np.frombuffer(c, dtype=np.uint8)
    return z, payload, offs


# ---------------------------------------------------------------------------
# result of encode
# ---------------------------------------------------------------------------

class Encoded:
    """Host-side view of everything F, T and C produce for one input."""

    def __init__(self, c: EncodedC):
        B = int(c.n_blocks)
        self.rle_mode = bool(c.rle_mode)
        self.n_blocks = B
        self.n = int(c.n)
        self.block_size = int(c.block_size)
        self.block_sizes = np.ctypeslib.as_array(c.block_sizes, (B,)).copy()
        self.sym_sizes = np.ctypeslib.as_array(c.sym_sizes, (B,)).copy()
        self.hist = np.ctypeslib.as_array(c.hist, (B, 256)).copy()
        self.hist_orig = np.ctypeslib.as_array(c.hist_orig, (B, 256)).copy() if c.hist_orig else None
        self.code_len = np.ctypeslib.as_array(c.code_len, (B, 256)).copy()
        self.code_bits = np.ctypeslib.as_array(c.code_bits, (B, 256)).copy()
        self.z = np.ctypeslib.as_array(c.z, (B,)).copy()
        self.payload_offsets = np.ctypeslib.as_array(c.payload_offsets, (B,)).copy()
        self.payload = np.ctypeslib.as_array(c.payload, (int(c.payload_total),)).copy() if c.payload_total else np.zeros(0, np.uint8)
        self.rle = (np.ctypeslib.as_array(c.rle, (int(c.rle_total),)).copy() if c.rle_mode and c.rle_total
                    else np.zeros(0, np.uint8))

    @property
    def mode(self) -> str:
        return "R" if self.rle_mode else "N"

    def freq_text(self) -> bytes:
        """`.rle.freq` in R mode, `.freq` in N mode."""
        return freq_text(self.mode, self.sym_sizes, self.hist)

    def freq_orig_text(self) -> bytes | None:
        """`.freq` written next to the R files by `-c f` (f.c:278-295,322-331)."""
        return None if self.hist_orig is None else freq_text("N", self.block_sizes, self.hist_orig)

    def cod_text(self) -> bytes:
        return cod_text(self.mode, self.sym_sizes, self.code_len, self.code_bits)

    def shaf_bytes(self) -> bytes:
        return shaf_bytes(self.z, self.payload, self.payload_offsets)

    def rle_bytes(self) -> bytes:
        return self.rle.tobytes()


# ---------------------------------------------------------------------------
# context
# ---------------------------------------------------------------------------

class DeviceBuffer:
    def __init__(self, sh: "Shafa", nbytes: int):
        self.sh, self.nbytes = sh, int(nbytes)
        self.ptr = sh.lib.shafa_cuda_dev_alloc(sh.ctx, max(16, (self.nbytes + 15) & ~15))
        if not self.ptr:
            raise ShafaError(2, sh.last_error())

    def upload(self, host: np.ndarray, offset: int = 0):
        host = np.ascontiguousarray(host)
        assert offset + host.nbytes <= max(16, (self.nbytes + 15) & ~15)
        self.sh._check(self.sh.lib.shafa_cuda_memcpy_h2d(self.sh.ctx, self.ptr + offset, host.ctypes.data, host.nbytes))
        self.sh.synchronize()
        return self

    def download(self, dtype=np.uint8, count: int | None = None, offset: int = 0) -> np.ndarray:
        item = np.dtype(dtype).itemsize
        count = (self.nbytes - offset) // item if count is None else count
        out = np.empty(count, dtype=dtype)
        if count:
            self.sh._check(self.sh.lib.shafa_cuda_memcpy_d2h(self.sh.ctx, out.ctypes.data, self.ptr + offset, out.nbytes))
            self.sh.synchronize()
        return out

    def free(self):
        if self.ptr:
            self.sh.lib.shafa_cuda_dev_free(self.sh.ctx, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Shafa:
    """One GPU context.  `encode`/`decode` are the whole-shard host-buffer calls;
    the *_blocks methods are the per-module device calls."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        if self.lib.shafa_cuda_device_count() <= device:
            raise ShafaError(7, "no CUDA device: libshafa_cuda has no CPU fallback")
        ctx = C.c_void_p()
        rc = self.lib.shafa_cuda_create(device, C.byref(ctx))
        if rc:
            raise ShafaError(rc, "shafa_cuda_create")
        self.ctx = ctx
        self.device = device

    # -- plumbing
    def close(self):
        if getattr(self, "ctx", None):
            self.lib.shafa_cuda_destroy(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def last_error(self) -> str:
        return self.lib.shafa_cuda_last_error(self.ctx).decode(errors="replace")

    def _check(self, rc: int):
        if rc:
            raise ShafaError(rc, self.last_error())

    def synchronize(self):
        self._check(self.lib.shafa_cuda_synchronize(self.ctx))

    def set_stream(self, stream_handle: int | None):
        self._check(self.lib.shafa_cuda_set_stream(self.ctx, C.c_void_p(stream_handle or 0)))

    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def to_device(self, host) -> DeviceBuffer:
        host = np.ascontiguousarray(host)
        return self.alloc(host.nbytes).upload(host)

    def profile(self, on: bool):
        self._check(self.lib.shafa_cuda_profile_enable(self.ctx, int(on)))

    def profile_reset(self):
        self._check(self.lib.shafa_cuda_profile_reset(self.ctx))

    def profile_read(self) -> dict:
        arr = (KernelStat * 64)()
        n = C.c_int(0)
        self._check(self.lib.shafa_cuda_profile_read(self.ctx, arr, 64, C.byref(n)))
        return {arr[i].name.decode(): {"launches": int(arr[i].launches), "ms": float(arr[i].total_ms)} for i in range(min(n.value, 64))}

    def launch_count(self) -> int:
        return int(self.lib.shafa_cuda_launch_count(self.ctx))

    # -- per-module device calls (pointers: DeviceBuffer or int)
    @staticmethod
    def _p(x):
        return x.ptr if isinstance(x, DeviceBuffer) else x

    def rle_gate(self, n0: int, rle0: int, force_rle: bool = False) -> bool:
        return bool(self.lib.shafa_cuda_rle_gate(n0, rle0, int(force_rle)))

    def rle_encode_blocks(self, d_in, n: int, block_size: int, d_rle=None, rle_capacity: int = 0, d_hist=None) -> np.ndarray:
        B = (n + block_size - 1) // block_size
        sizes = np.zeros(B, dtype=np.uint64)
        self._check(self.lib.shafa_cuda_rle_encode_blocks(self.ctx, self._p(d_in), n, block_size, self._p(d_rle) if d_rle is not None else None,
                                                         rle_capacity, sizes.ctypes.data, self._p(d_hist) if d_hist is not None else None))
        return sizes

    def histogram_blocks(self, d_in, offsets, d_hist):
        offsets = _u64(offsets)
        self._check(self.lib.shafa_cuda_histogram_blocks(self.ctx, self._p(d_in), offsets.ctypes.data, len(offsets) - 1, self._p(d_hist)))

    def sf_build_codes(self, d_hist, n_blocks: int, d_len, d_code):
        self._check(self.lib.shafa_cuda_sf_build_codes(self.ctx, self._p(d_hist), n_blocks, self._p(d_len), self._p(d_code)))

    def sf_encode_blocks(self, d_sym, offsets, d_len, d_code, d_hist, d_payload, payload_capacity: int):
        offsets = _u64(offsets)
        B = len(offsets) - 1
        poff = np.zeros(B, dtype=np.uint64)
        z = np.zeros(B, dtype=np.uint64)
        self._check(self.lib.shafa_cuda_sf_encode_blocks(self.ctx, self._p(d_sym), offsets.ctypes.data, B, self._p(d_len), self._p(d_code),
                                                        self._p(d_hist) if d_hist is not None else None, self._p(d_payload),
                                                        payload_capacity, poff.ctypes.data, z.ctypes.data))
        return poff, z

    def sf_decode_blocks(self, d_payload, payload_offsets, z, n, d_len, d_code, d_sym, sym_capacity: int):
        payload_offsets, z, n = _u64(payload_offsets), _u64(z), _u64(n)
        self._check(self.lib.shafa_cuda_sf_decode_blocks(self.ctx, self._p(d_payload), payload_offsets.ctypes.data, z.ctypes.data,
                                                        n.ctypes.data, len(n), self._p(d_len), self._p(d_code), self._p(d_sym), sym_capacity))

    def rle_decode_blocks(self, d_rle, offsets, d_out, out_capacity: int) -> np.ndarray:
        offsets = _u64(offsets)
        fin = np.zeros(len(offsets) - 1, dtype=np.uint64)
        self._check(self.lib.shafa_cuda_rle_decode_blocks(self.ctx, self._p(d_rle), offsets.ctypes.data, len(offsets) - 1,
                                                         self._p(d_out), out_capacity, fin.ctypes.data))
        return fin

    # -- whole-shard host calls
    def encode_raw(self, data, block_size: int, force_rle=False, force_freq=False, rle_buf=None, payload_buf=None) -> EncodedC:
        """F -> T -> C; returns the raw C struct (caller frees with encoded_free)."""
        a = _u8(data)
        c = EncodedC()
        if rle_buf is not None:
            c.rle = C.cast(rle_buf.ctypes.data, C.POINTER(C.c_uint8))
            c.rle_capacity = rle_buf.nbytes
        if payload_buf is not None:
            c.payload = C.cast(payload_buf.ctypes.data, C.POINTER(C.c_uint8))
            c.payload_capacity = payload_buf.nbytes
        self._check(self.lib.shafa_cuda_encode_host(self.ctx, a.ctypes.data, a.size, block_size, int(force_rle), int(force_freq), C.byref(c)))
        return c

    def encoded_free(self, c: EncodedC):
        self.lib.shafa_cuda_encoded_free(C.byref(c))

    def encode(self, data, block_size: int, force_rle=False, force_freq=False) -> Encoded:
        c = self.encode_raw(data, block_size, force_rle, force_freq)
        try:
            return Encoded(c)
        finally:
            self.encoded_free(c)

    def encode_shard(self, data, block_size: int, rle_mode: bool, force_freq=False) -> Encoded:
        """F -> T -> C for a shard that does not hold the file's block 0 (mode decided by its owner)."""
        a = _u8(data)
        c = EncodedC()
        self._check(self.lib.shafa_cuda_encode_shard_host(self.ctx, a.ctypes.data, a.size, block_size, int(rle_mode), int(force_freq), C.byref(c)))
        try:
            return Encoded(c)
        finally:
            self.encoded_free(c)

    def sf_encode_wide_host(self, sym, offsets, code_len, code4):
        """Tables with codes of 65..255 bits (c.c:18): -> (payload with 16-byte aligned slots, payload_offsets[B], z[B])."""
        sym, code_len = _u8(sym), _u8(code_len)
        offsets, code4 = _u64(offsets), _u64(code4)
        nb = len(offsets) - 1
        poff, z = np.zeros(nb, dtype=np.uint64), np.zeros(nb, dtype=np.uint64)
        total = C.c_uint64(0)
        cap = int(offsets[-1] - offsets[0]) * 2 + 64 * nb + 4096
        for _ in range(2):
            payload = np.zeros(cap, dtype=np.uint8)
            rc = self.lib.shafa_cuda_sf_encode_wide_host(self.ctx, sym.ctypes.data, offsets.ctypes.data, nb, code_len.ctypes.data, code4.ctypes.data,
                                                        payload.ctypes.data, cap, poff.ctypes.data, z.ctypes.data, C.byref(total))
            if rc != 2 or total.value <= cap:
                break
            cap = total.value
        self._check(rc)
        return payload[: total.value], poff, z

    def sf_decode_wide_host(self, payload, payload_offsets, z, n, code_len, code4):
        payload, code_len = _u8(payload), _u8(code_len)
        payload_offsets, z, n, code4 = _u64(payload_offsets), _u64(z), _u64(n), _u64(code4)
        out = np.empty(int(n.sum()) + 64, dtype=np.uint8)
        self._check(self.lib.shafa_cuda_sf_decode_wide_host(self.ctx, payload.ctypes.data, payload_offsets.ctypes.data, z.ctypes.data, n.ctypes.data,
                                                           len(n), code_len.ctypes.data, code4.ctypes.data, out.ctypes.data, out.nbytes))
        return out[: int(n.sum())]

    def decode_tables(self, payload, payload_offsets, z, n, code_len, code_bits, rle_expand: bool, out_capacity: int,
                      out: np.ndarray | None = None):
        payload, code_len = _u8(payload), _u8(code_len)
        payload_offsets, z, n, code_bits = _u64(payload_offsets), _u64(z), _u64(n), _u64(code_bits)
        if out is None:
            out = np.empty(out_capacity, dtype=np.uint8)
        fin = np.zeros(len(n), dtype=np.uint64)
        size = C.c_uint64(0)
        self._check(self.lib.shafa_cuda_decode_host(self.ctx, payload.ctypes.data, payload_offsets.ctypes.data, z.ctypes.data, n.ctypes.data,
                                                   len(n), code_len.ctypes.data, code_bits.ctypes.data, int(rle_expand),
                                                   out.ctypes.data, out.nbytes, C.byref(size), fin.ctypes.data))
        return out[: size.value], fin

    def decode(self, enc: Encoded, rle_expand: bool | None = None) -> np.ndarray:
        """D: SF decode (+ RLE expansion when the stream is RLE and rle_expand is not False)."""
        expand = enc.rle_mode if rle_expand is None else (rle_expand and enc.rle_mode)
        cap = enc.n if expand else int(enc.sym_sizes.sum())
        out, _ = self.decode_tables(enc.payload, enc.payload_offsets, enc.z, enc.sym_sizes, enc.code_len, enc.code_bits,
                                    expand, cap + 64)
        return out

    def rle_decode_host(self, rle, offsets, out_capacity: int):
        rle, offsets = _u8(rle), _u64(offsets)
        out = np.empty(out_capacity, dtype=np.uint8)
        fin = np.zeros(len(offsets) - 1, dtype=np.uint64)
        size = C.c_uint64(0)
        self._check(self.lib.shafa_cuda_rle_decode_host(self.ctx, rle.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                                                       out.ctypes.data, out.nbytes, C.byref(size), fin.ctypes.data))
        return out[: size.value], fin


class ShafaMulti:
    """ONE file sharded over several GPUs by this one process (shafa_cuda_*_multi): chunks of whole
    blocks go to whichever device is free, results land in block order in host buffers.  No collective.  The same
    device may be listed more than once (useful on a one-GPU box: the sharding logic is the same)."""

    def __init__(self, devices):
        self.lib = load_library()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self.lib.shafa_cuda_multi_create(devs, len(devices), C.byref(h))
        if rc:
            raise ShafaError(rc, f"shafa_cuda_multi_create({list(devices)})")
        self.h = h
        self.devices = list(devices)

    def close(self):
        if getattr(self, "h", None):
            self.lib.shafa_cuda_multi_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def last_error(self) -> str:
        return self.lib.shafa_cuda_multi_last_error(self.h).decode(errors="replace")

    def launch_count(self) -> int:
        return int(self.lib.shafa_cuda_multi_launch_count(self.h))

    def _check(self, rc: int):
        if rc:
            raise ShafaError(rc, self.last_error())

    def encode_raw(self, data, block_size: int, force_rle=False, force_freq=False, rle_buf=None, payload_buf=None) -> EncodedC:
        a = _u8(data)
        c = EncodedC()
        if rle_buf is not None:
            c.rle = C.cast(rle_buf.ctypes.data, C.POINTER(C.c_uint8))
            c.rle_capacity = rle_buf.nbytes
        if payload_buf is not None:
            c.payload = C.cast(payload_buf.ctypes.data, C.POINTER(C.c_uint8))
            c.payload_capacity = payload_buf.nbytes
        self._check(self.lib.shafa_cuda_encode_host_multi(self.h, a.ctypes.data, a.size, block_size, int(force_rle), int(force_freq), C.byref(c)))
        return c

    def encoded_free(self, c: EncodedC):
        self.lib.shafa_cuda_encoded_free(C.byref(c))

    def encode(self, data, block_size: int, force_rle=False, force_freq=False) -> Encoded:
        c = self.encode_raw(data, block_size, force_rle, force_freq)
        try:
            return Encoded(c)
        finally:
            self.encoded_free(c)

    def decode_tables(self, payload, payload_offsets, z, n, code_len, code_bits, rle_expand: bool, out_capacity: int,
                      out: np.ndarray | None = None):
        payload, code_len = _u8(payload), _u8(code_len)
        payload_offsets, z, n, code_bits = _u64(payload_offsets), _u64(z), _u64(n), _u64(code_bits)
        if out is None:
            out = np.empty(out_capacity, dtype=np.uint8)
        fin = np.zeros(len(n), dtype=np.uint64)
        size = C.c_uint64(0)
        self._check(self.lib.shafa_cuda_decode_host_multi(self.h, payload.ctypes.data, payload_offsets.ctypes.data, z.ctypes.data, n.ctypes.data,
                                                         len(n), code_len.ctypes.data, code_bits.ctypes.data, int(rle_expand),
                                                         out.ctypes.data, out.nbytes, C.byref(size), fin.ctypes.data))
        return out[: size.value], fin

    def decode(self, enc: Encoded, rle_expand: bool | None = None) -> np.ndarray:
        expand = enc.rle_mode if rle_expand is None else (rle_expand and enc.rle_mode)
        cap = enc.n if expand else int(enc.sym_sizes.sum())
        out, _ = self.decode_tables(enc.payload, enc.payload_offsets, enc.z, enc.sym_sizes, enc.code_len, enc.code_bits, expand, cap + 64)
        return out

    def rle_decode_host(self, rle, offsets, out_capacity: int, out: np.ndarray | None = None):
        rle, offsets = _u8(rle), _u64(offsets)
        if out is None:
            out = np.empty(out_capacity, dtype=np.uint8)
        fin = np.zeros(len(offsets) - 1, dtype=np.uint64)
        size = C.c_uint64(0)
        self._check(self.lib.shafa_cuda_rle_decode_host_multi(self.h, rle.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                                                             out.ctypes.data, out.nbytes, C.byref(size), fin.ctypes.data))
        return out[: size.value], fin
