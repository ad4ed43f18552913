"""Minimal reader for R's `save()` files (.rda / .RData, serialization format 2/3, XDR, gzip/bzip2/xz or plain).

Host-side stand-in for `load()` in `ReadModel` (/root/reference/src/SAIGE/R/readInGLMM.R:39-45): step 2 consumes the
null model that step 1 wrote with `save(modglmm, file = ...)` (FG.R:1297-1301).  Only the SEXP types that occur in
SAIGE model files are supported: NULL, symbols, pairlists, lists, character / logical / integer / real vectors,
attributes (names, dim, dimnames, class), reference objects and the ALTREP compact sequences / wrappers.

Returns plain Python objects: named lists -> dict (insertion ordered), unnamed lists -> list, atomic vectors ->
numpy arrays (matrices reshaped column-major via `dim`), length-1 atomics stay arrays; NA_integer_ -> masked as -2^31.
"""
import bz2
import gzip
import lzma
import struct

import numpy as np

NILVALUE_SXP, GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP = 254, 253, 242, 241
REFSXP, PERSISTSXP, PACKAGESXP, NAMESPACESXP, BASENAMESPACE_SXP, MISSINGARG_SXP, UNBOUNDVALUE_SXP = 255, 247, 250, 249, 247, 251, 252
ALTREP_SXP, ATTRLISTSXP, ATTRLANGSXP = 238, 239, 240
BCODESXP, BCREPDEF, BCREPREF = 21, 244, 243
NILSXP, SYMSXP, LISTSXP, CLOSXP, ENVSXP, PROMSXP, LANGSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, CPLXSXP, STRSXP, VECSXP, EXPRSXP, RAWSXP, S4SXP = (
    0, 1, 2, 3, 4, 5, 6, 9, 10, 13, 14, 15, 16, 19, 20, 24, 25)


class RObject:
    """Anything we do not convert (closures, environments, language objects): kept opaque."""

    def __init__(self, kind):
        self.kind = kind

    def __repr__(self):
        return "<R %s>" % self.kind


class RList(dict):
    """A named R list that carries an S3 class (e.g. obj.noK: class "SA_NULL", FG.R:589)."""

    def __init__(self, *a, r_class=None, **kw):
        super().__init__(*a, **kw)
        self.r_class = r_class


class _Reader:
    def __init__(self, data):
        self.b, self.p, self.refs = data, 0, []

    def int(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def length(self):
        n = self.int()
        if n == -1:
            hi, lo = struct.unpack_from(">II", self.b, self.p)
            self.p += 8
            n = (hi << 32) | lo
        return n

    def bytes(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.int()
        typ = flags & 0xFF
        has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if typ in (NILVALUE_SXP, NILSXP):
            return None
        if typ in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, MISSINGARG_SXP, UNBOUNDVALUE_SXP, 247):
            return RObject("env/special %d" % typ)
        if typ == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if typ == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if typ in (PACKAGESXP, NAMESPACESXP, PERSISTSXP):
            self.int()          # 0
            n = self.int()
            obj = RObject("namespace " + " ".join(str(self.item()) for _ in range(n)))
            self.refs.append(obj)
            return obj
        if typ == ENVSXP:
            obj = RObject("environment")
            self.refs.append(obj)
            self.int()          # locked
            for _ in range(4):  # enclos, frame, hashtab, attrib
                self.item()
            return obj
        if typ in (LISTSXP, LANGSXP, CLOSXP, PROMSXP, ATTRLISTSXP, ATTRLANGSXP):
            # pairlist chain: iterate instead of recursing on the tail
            out, first = [], True
            while True:
                if not first:
                    flags = self.int()
                    typ = flags & 0xFF
                    has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
                    if typ in (NILVALUE_SXP, NILSXP):
                        break
                    if typ not in (LISTSXP, LANGSXP, CLOSXP, PROMSXP, ATTRLISTSXP, ATTRLANGSXP):
                        self.p -= 4
                        out.append((None, self.item()))
                        break
                first = False
                if has_attr:
                    self.item()
                tag = self.item() if has_tag else None
                out.append((tag, self.item()))
            return out
        if typ == CHARSXP:
            n = self.int()
            return None if n == -1 else self.bytes(n).decode("utf-8", "replace")
        attr = None
        if typ == ALTREP_SXP:
            info, state, attr_ = self.item(), self.item(), self.item()
            cls = info[0][1] if isinstance(info, list) else str(info)
            val = self._altrep(str(cls), state)
            return self._with_attr(val, attr_)
        if typ == LGLSXP or typ == INTSXP:
            n = self.length()
            val = np.frombuffer(self.bytes(4 * n), dtype=">i4").astype(np.int32)
            if typ == LGLSXP:
                val = np.where(val == -2147483648, -1, val).astype(np.int8)
        elif typ == REALSXP:
            n = self.length()
            val = np.frombuffer(self.bytes(8 * n), dtype=">f8").astype(np.float64)
        elif typ == CPLXSXP:
            n = self.length()
            val = np.frombuffer(self.bytes(16 * n), dtype=">c16").astype(np.complex128)
        elif typ == STRSXP:
            n = self.length()
            val = [self.item() for _ in range(n)]
        elif typ in (VECSXP, EXPRSXP):
            n = self.length()
            val = [self.item() for _ in range(n)]
        elif typ == RAWSXP:
            n = self.length()
            val = np.frombuffer(self.bytes(n), dtype=np.uint8).copy()
        elif typ == S4SXP:
            val = RObject("S4")
        elif typ in (7, 8):      # SPECIALSXP / BUILTINSXP
            n = self.int()
            val = RObject("builtin " + self.bytes(n).decode())
        elif typ == 22:          # EXTPTRSXP: a reference object holding (protected, tag)
            val = RObject("externalptr")
            self.refs.append(val)
            self.item()
            self.item()
        elif typ == 23:          # WEAKREFSXP
            val = RObject("weakref")
            self.refs.append(val)
        elif typ == BCODESXP:    # byte-compiled closure bodies (older model files keep family functions): parsed, not kept
            self._bc_reps = [None] * self.int()
            self._bc1()
            val = RObject("bytecode")
        else:
            raise ValueError("unsupported SEXP type %d at byte %d" % (typ, self.p))
        if has_attr:
            attr = self.item()
        return self._with_attr(val, attr)

    def _bc1(self):
        # ReadBC1 (R serialize.c): the code vector, then the constant pool
        self.item()
        for _ in range(self.int()):
            t = self.int()
            if t == BCODESXP:
                self._bc1()
            elif t in (LANGSXP, LISTSXP, BCREPDEF, BCREPREF, ATTRLANGSXP, ATTRLISTSXP):
                self._bclang(t)
            else:
                self.item()          # the type word is a prefix; a complete item follows

    def _bclang(self, t):
        # ReadBCLang: language objects of the constant pool, with back references between them
        if t == BCREPREF:
            self.int()
            return
        if t not in (BCREPDEF, LANGSXP, LISTSXP, ATTRLANGSXP, ATTRLISTSXP):
            self.item()
            return
        if t == BCREPDEF:
            self.int()               # position in the reps table
            t = self.int()
        if t in (ATTRLANGSXP, ATTRLISTSXP):
            self.item()              # attributes
        self.item()                  # tag
        self._bclang(self.int())     # car
        self._bclang(self.int())     # cdr

    def _altrep(self, cls, state):
        if cls in ("compact_intseq", "compact_realseq"):
            n, start, step = (float(x) for x in state[:3])
            arr = start + step * np.arange(int(n))
            return arr.astype(np.int32 if cls == "compact_intseq" else np.float64)
        if cls.startswith("wrap_"):
            return state[0][1] if isinstance(state, list) and state and isinstance(state[0], tuple) else state
        if cls == "deferred_string":
            src = state[0][1] if isinstance(state[0], tuple) else state[0]
            return [("%d" % v if float(v).is_integer() else repr(float(v))) for v in np.asarray(src)]
        raise ValueError("unsupported ALTREP class " + cls)

    @staticmethod
    def _with_attr(val, attr):
        if not attr:
            return val
        a = {str(k): v for k, v in attr}
        if isinstance(val, np.ndarray) and "dim" in a:
            val = val.reshape(tuple(int(d) for d in a["dim"]), order="F")
        if isinstance(val, list) and "names" in a and a["names"] is not None:
            names = list(a["names"])
            if len(names) == len(val) and all(n not in (None, "") for n in names):
                if a.get("class"):
                    return RList(zip(names, val), r_class=list(a["class"]))
                return dict(zip(names, val))
        return val


def load_rda(path):
    """Returns {object name: value} for every object saved in the file."""
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    if raw[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError("%s: not an R save() file (magic %r)" % (path, raw[:5]))
    r = _Reader(raw)
    r.p = 5
    if r.bytes(2) != b"X\n":
        raise ValueError("only XDR serialization is supported")
    version = r.int()
    r.int(); r.int()             # writer version, min reader version
    if version == 3:
        n = r.int()
        r.bytes(n)               # native encoding
    top = r.item()
    return {str(k): v for k, v in top}


# ---------------------------------------------------------------------------------------------------------------------
# writer: the counterpart of R's save(modglmm, file = modelOut) (FG.R:1297-1301), so that a null model fitted through the
# Python mirror of fitNULLGLMM can be consumed by load() in the reference's step 2 (R/readInGLMM.R:39-45)
# ---------------------------------------------------------------------------------------------------------------------
class _Writer:
    """Serialization format version 2 (what the reference's bundled .rda files use; every R >= 1.4 reads it), XDR."""

    def __init__(self):
        self.out, self.sym = [], {}

    def int(self, v):
        self.out.append(struct.pack(">i", v))

    def length(self, n):
        if n > 2147483647:
            self.int(-1)
            self.out.append(struct.pack(">II", n >> 32, n & 0xFFFFFFFF))
        else:
            self.int(n)

    def charsxp(self, s):
        if s is None:
            self.int(CHARSXP)              # NA_character_
            self.int(-1)
            return
        b = s.encode("utf-8")
        levels = 64 if all(c < 128 for c in b) else 8          # ASCII_MASK / UTF8_MASK in the gp field
        self.int(CHARSXP | (levels << 12))
        self.int(len(b))
        self.out.append(b)

    def symbol(self, name):
        if name in self.sym:                                   # later occurrences are back references
            self.int((self.sym[name] << 8) | REFSXP)
            return
        self.sym[name] = len(self.sym) + 1
        self.int(SYMSXP)
        self.charsxp(name)

    def attributes(self, attr):
        for k, v in attr:
            self.int(LISTSXP | 0x400)                          # pairlist node with a tag
            self.symbol(k)
            self.item(v)
        self.int(NILVALUE_SXP)

    def vector_header(self, typ, n, attr):
        self.int(typ | (0x200 if attr else 0))
        self.length(n)

    def item(self, v, attr=None):
        attr = list(attr) if attr else []
        if v is None:
            self.int(NILVALUE_SXP)
            return
        if isinstance(v, dict):
            keys = list(v.keys())
            cls = getattr(v, "r_class", None)
            self.int(VECSXP | 0x200 | (0x100 if cls else 0))       # 0x100: is an object (has a class attribute)
            self.length(len(keys))
            for k in keys:
                self.item(v[k])
            self.attributes([("names", [str(k) for k in keys])] + ([("class", list(cls))] if cls else []) + attr)
            return
        if isinstance(v, (bool, np.bool_)):
            v = np.array([v], dtype=np.bool_)
        elif isinstance(v, (int, np.integer)):
            v = np.array([v], dtype=np.int32)
        elif isinstance(v, (float, np.floating)):
            v = np.array([v], dtype=np.float64)
        elif isinstance(v, str):
            v = [v]
        if isinstance(v, (list, tuple)):
            if any(isinstance(x, str) for x in v) and all(isinstance(x, str) or x is None for x in v):      # [None] is list(NULL)
                self.vector_header(STRSXP, len(v), attr)
                for x in v:
                    self.charsxp(x)
            else:
                self.vector_header(VECSXP, len(v), attr)
                for x in v:
                    self.item(x)
            if attr:
                self.attributes(attr)
            return
        if isinstance(v, np.ndarray):
            if v.ndim >= 2:
                attr = [("dim", np.array(v.shape, dtype=np.int32))] + attr
            flat = np.asarray(v).reshape(-1, order="F")
            if v.dtype == np.bool_ or v.dtype == np.int8:          # the reader returns logicals as int8 (NA = -1)
                typ, data = LGLSXP, np.where(flat.astype(np.int64) < 0, -2147483648, flat.astype(np.int64)).astype(">i4")
            elif np.issubdtype(v.dtype, np.integer):
                typ, data = INTSXP, flat.astype(">i4")
            elif np.issubdtype(v.dtype, np.floating):
                typ, data = REALSXP, flat.astype(">f8")
            else:
                raise TypeError("cannot serialize an array of dtype %s" % v.dtype)
            self.vector_header(typ, flat.size, attr)
            self.out.append(data.tobytes())
            if attr:
                self.attributes(attr)
            return
        raise TypeError("cannot serialize %r" % type(v))


def save_rda(path, objects, compress=True):
    """save(<objects>, file = path): `objects` maps R object names to values.  dict -> named list, list of str -> character
    vector, other list -> unnamed list, numpy arrays -> logical (bool / int8) / integer / double vectors and matrices
    (column-major, `dim` attribute), scalars -> length-1 vectors, None -> NULL.  gzip-compressed like R's default."""
    w = _Writer()
    w.out.append(b"RDX2\nX\n")
    w.int(2)
    w.int(0x00030603)            # written by R 3.6.3 (any version R accepts); readable from R 2.3.0
    w.int(0x00020300)
    for name, v in objects.items():
        w.int(LISTSXP | 0x400)
        w.symbol(str(name))
        w.item(v)
    w.int(NILVALUE_SXP)
    raw = b"".join(w.out)
    with open(path, "wb") as f:
        f.write(gzip.compress(raw, 6, mtime=0) if compress else raw)
