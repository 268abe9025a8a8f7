"""Minimal reader for R's XDR serialisation (the .rda files of the reference's data/ directory),
enough for lists of numeric / integer / logical / character vectors with attributes and simple S4
objects.  Written for this repository (R is not installed); format as documented in "R Internals",
section 1.8 Serialization Formats.  Used by make_scotcancer.py only."""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct


class _Reader:
    def __init__(self, data: bytes):
        self.d = data
        self.i = 0
        self.refs = []

    def int(self) -> int:
        v = struct.unpack_from(">i", self.d, self.i)[0]
        self.i += 4
        return v

    def dbl(self) -> float:
        v = struct.unpack_from(">d", self.d, self.i)[0]
        self.i += 8
        return v

    def raw(self, n: int) -> bytes:
        v = self.d[self.i:self.i + n]
        self.i += n
        return v

    def length(self) -> int:
        n = self.int()
        if n == -1:
            hi, lo = self.int(), self.int()
            n = (hi << 32) + lo
        return n

    def item(self):
        flags = self.int()
        typ = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if typ == 254:      # NILVALUE_SXP
            return None
        if typ == 253:      # GLOBALENV
            return "<globalenv>"
        if typ == 242:      # EMPTYENV
            return "<emptyenv>"
        if typ == 255:      # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if typ == 1:        # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if typ in (2, 6):   # LISTSXP / LANGSXP: pairlist -> list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                nflags_pos = self.i
                nflags = self.int()
                ntyp = nflags & 0xFF
                if ntyp == 254:
                    return out
                if ntyp not in (2, 6):
                    raise ValueError(f"unexpected pairlist tail type {ntyp}")
                has_attr = bool(nflags & 0x200)
                has_tag = bool(nflags & 0x400)
                del attr, nflags_pos
        if typ == 9:        # CHARSXP
            n = self.int()
            return None if n == -1 else self.raw(n).decode("utf-8", "replace")
        obj = None
        if typ == 10 or typ == 13:   # LGLSXP / INTSXP
            n = self.length()
            obj = list(struct.unpack_from(f">{n}i", self.d, self.i))
            self.i += 4 * n
        elif typ == 14:              # REALSXP
            n = self.length()
            obj = list(struct.unpack_from(f">{n}d", self.d, self.i))
            self.i += 8 * n
        elif typ == 16:              # STRSXP
            n = self.length()
            obj = [self.item() for _ in range(n)]
        elif typ == 19 or typ == 20:  # VECSXP / EXPRSXP
            n = self.length()
            obj = [self.item() for _ in range(n)]
        elif typ == 25:              # S4SXP: slots live in the attributes
            obj = {}
        else:
            raise ValueError(f"unsupported SEXP type {typ} at byte {self.i}")
        attrs = dict((k, v) for k, v in self.item()) if has_attr else {}
        return {"value": obj, "attr": attrs} if attrs else obj


def load_rda(path: str):
    """Returns the list of (name, object) pairs stored in an .rda file."""
    raw = open(path, "rb").read()
    for dec in (bz2.decompress, gzip.decompress, lzma.decompress, lambda b: b):
        try:
            data = dec(raw)
            break
        except Exception:
            continue
    if not data.startswith(b"RDX2\nX\n") and not data.startswith(b"RDX3\nX\n"):
        raise ValueError("not an XDR .rda file")
    r = _Reader(data[7:])
    version = r.int()
    r.int()
    r.int()
    if version == 3:
        n = r.int()
        r.raw(n)
    return r.item()
