"""Reader for R's `.rda` / `.RData` files (the container of the reference's dataset, data/usmacro_growth.rda).

`save()` writes a (bzip2 | gzip | xz)-compressed stream that starts with the magic line ``RDX2`` / ``RDX3`` followed by
R's XDR serialisation (src/main/serialize.c of R, format versions 2 and 3): big-endian, one header word per
object whose low byte is the SEXP type and whose bits 8-10 say "is an object / has attributes / has a tag".
Only what data matrices need is decoded: pairlists, symbols, character / integer / logical / real / string /
generic vectors, NULL, and back-references to symbols.  A numeric matrix comes back as `RMatrix` (values in
column-major order as R holds them, plus dimnames).

    objs = read_rda("usmacro_growth.rda");  m = objs["usmacro_growth"];  m.values.shape == (247, 21)
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

NILVALUE_SXP, GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, MISSINGARG_SXP, UNBOUNDVALUE_SXP = 254, 253, 242, 241, 251, 252
REFSXP, NAMESPACESXP, PACKAGESXP, PERSISTSXP = 255, 249, 250, 247
ALTREP_SXP, ATTRLISTSXP, ATTRLANGSXP = 238, 239, 240
NILSXP, SYMSXP, LISTSXP, CLOSXP, ENVSXP, PROMSXP, LANGSXP = 0, 1, 2, 3, 4, 5, 6
CHARSXP, LGLSXP, INTSXP, REALSXP, CPLXSXP, STRSXP, VECSXP, EXPRSXP, RAWSXP = 9, 10, 13, 14, 15, 16, 19, 20, 24
NA_INT = -2147483648


@dataclass
class RMatrix:
    values: np.ndarray                      # (nrow, ncol) float64 / int32, Fortran order like R
    rownames: list | None = None
    colnames: list | None = None
    attributes: dict = field(default_factory=dict)

    def columns(self, names):
        """x[, c("a", "b")]"""
        idx = [self.colnames.index(n) for n in names]
        return self.values[:, idx]


class _Reader:
    def __init__(self, buf: bytes):
        self.b, self.o, self.refs = buf, 0, []

    def take(self, n):
        v = self.b[self.o:self.o + n]
        if len(v) != n:
            raise ValueError("truncated R serialisation stream")
        self.o += n
        return v

    def i32(self):
        return struct.unpack(">i", self.take(4))[0]

    def length(self):
        n = self.i32()
        if n == -1:                          # long vector: two more words
            hi, lo = struct.unpack(">II", self.take(8))
            n = (hi << 32) | lo
        return n

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        is_obj, has_attr, has_tag = bool(flags & (1 << 8)), bool(flags & (1 << 9)), bool(flags & (1 << 10))
        if t == NILVALUE_SXP or t == NILSXP:
            return None
        if t in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, MISSINGARG_SXP, UNBOUNDVALUE_SXP):
            return None
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            self.i32()
            n = self.i32()
            v = [self.item() for _ in range(n)]
            self.refs.append(v)
            return v
        if t in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP):
            # pairlist: (attributes)? (tag)? car cdr  - walked iteratively, returned as a list of (tag, value)
            out = []
            while True:
                attrs = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                del attrs
                flags = self.i32()
                t2 = flags & 0xFF
                if t2 in (NILVALUE_SXP, NILSXP):
                    break
                if t2 not in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP):
                    raise ValueError(f"unsupported pairlist tail of type {t2}")
                has_attr, has_tag = bool(flags & (1 << 9)), bool(flags & (1 << 10))
            return out
        if t == CHARSXP:
            n = self.i32()
            return None if n == -1 else self.take(n).decode("utf-8", "replace")
        if t == LGLSXP or t == INTSXP:
            n = self.length()
            v = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
        elif t == REALSXP:
            n = self.length()
            v = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif t == STRSXP:
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t in (VECSXP, EXPRSXP):
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t == RAWSXP:
            n = self.length()
            v = self.take(n)
        else:
            raise ValueError(f"R object of SEXP type {t} is not supported by this reader")
        attrs = dict((k, a) for k, a in self.item()) if has_attr else {}
        del is_obj
        return self._with_attributes(v, attrs)

    @staticmethod
    def _with_attributes(v, attrs):
        if not attrs:
            return v
        if isinstance(v, np.ndarray) and "dim" in attrs and len(attrs["dim"]) == 2:
            nr, nc = (int(x) for x in attrs["dim"])
            dn = attrs.get("dimnames") or [None, None]
            rest = {k: a for k, a in attrs.items() if k not in ("dim", "dimnames")}
            return RMatrix(np.asfortranarray(v.reshape((nr, nc), order="F")), dn[0], dn[1], rest)
        if isinstance(v, list) and "names" in attrs:
            return dict(zip(attrs["names"], v))
        return v


def _decompress(raw: bytes) -> bytes:
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def read_rda(path) -> dict:
    """All objects of an `.rda` file as {name: value}."""
    buf = _decompress(Path(path).read_bytes())
    if buf[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError("not an R data file written by save() (magic RDX2 / RDX3 missing)")
    if buf[5:7] != b"X\n":
        raise ValueError("only the XDR serialisation is supported")
    r = _Reader(buf)
    r.o = 7
    version = r.i32()
    r.i32()   # R version that wrote the file
    r.i32()   # minimal R version able to read it
    if version == 3:
        r.take(r.i32())                      # native encoding name
    elif version != 2:
        raise ValueError(f"serialisation format version {version} is not supported")
    top = r.item()
    if not isinstance(top, list):
        raise ValueError("an .rda file holds a pairlist of named objects")
    return {tag: val for tag, val in top}
