"""Step-2 genotype inputs other than PLINK: VCF (GT or DS field) and BGEN v1.2 layout 2, decoded on the host into the
dosage matrix `sgb_step2_test_dosages` consumes (copies of the tested allele per sample, doubles, negative = missing).

Host-side mirror of the reference's readers (/root/reference/src/SAIGE/src/VCF.cpp:120-256 `VcfClass::getOneMarker`, which
goes through the savvy library, and src/BGEN.cpp:132-345 `BgenClass::Parse2`, :360-520 `getOneMarker`):
  * VCF, GT: dosage = number of ALT alleles of the call, missing when an allele is `.`; DS: the number itself, `.` missing.
    Allele1 = REF, Allele2 = ALT (the tested allele).
  * BGEN: zlib-compressed probability blocks of unphased diploid biallelic variants, 8 or 16 bits per probability;
    P(AA), P(AB) of the FIRST allele A are stored, dosage of the first allele = 2 P(AA) + P(AB) (BGEN.cpp:218-223).
    AlleleOrder "ref-first" (the reader's default, BGEN.cpp:485-486): first allele = REF, tested allele = second;
    "alt-first": the first allele is the tested one.  Missing samples carry ploidy byte bit 7.
Both yield chunks (info rows, dosage matrix) so a scan never holds the whole file.  Only formats; no statistics here."""
import gzip
import struct
import warnings
import zlib

import numpy as np


def _open_text(path):
    with open(path, "rb") as f:
        gz = f.read(2) == b"\x1f\x8b"
    return gzip.open(path, "rt") if gz else open(path, "rt")


def vcf_samples(path):
    with _open_text(path) as f:
        for line in f:
            if line.startswith("#CHROM"):
                return line.rstrip("\n").split("\t")[9:]
    raise ValueError("%s: no #CHROM header line" % path)


def iter_vcf(path, field="DS", chunk=1000, only=None):
    """Yields (info, D): info = list of (CHR, POS, ID, REF, ALT), D = len(info) x n_samples doubles (negative = missing).
    only: a set of "chr:pos:ref:alt" keys -- other records are passed over after their first five fields."""
    if field not in ("GT", "DS"):
        raise ValueError("vcfField should be 'DS' or 'GT'")
    info, rows = [], []
    gt_lut = {}
    with _open_text(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            if only is not None:
                h5 = line.split("\t", 5)
                if len(h5) < 6 or "%s:%s:%s:%s" % (h5[0], h5[1], h5[3], h5[4]) not in only:
                    continue
            t = line.rstrip("\n").split("\t", 9)            # fixed fields + the sample part as one string
            if len(t) < 10:
                raise ValueError("%s: record %s has no sample columns" % (path, t[2] if len(t) > 2 else "?"))
            if "," in t[4]:
                raise NotImplementedError("%s: multi-allelic record %s (split it into biallelic records first)" % (path, t[2]))
            fmt = t[8].split(":")
            if field not in fmt:
                raise ValueError("%s: record %s has no %s field" % (path, t[2], field))
            k = fmt.index(field)
            row = None
            if len(fmt) == 1:
                row = _gt_fast(t[9]) if field == "GT" else _ds_fast(t[9])
            if row is None:
                cells = t[9].split("\t")
                vals = cells if len(fmt) == 1 else [x.split(":")[k] if x.count(":") >= k else "." for x in cells]
                if field == "GT":
                    row = np.empty(len(vals))
                    for i, v in enumerate(vals):
                        d = gt_lut.get(v)
                        if d is None:
                            al = v.replace("|", "/").split("/")
                            d = -1.0 if "." in al else float(sum(a != "0" for a in al))
                            gt_lut[v] = d
                        row[i] = d
                else:
                    row = np.array([-1.0 if v in (".", "") else float(v) for v in vals])
            info.append((t[0], t[1], t[2], t[3], t[4]))
            rows.append(row)
            if len(rows) == chunk:
                yield info, np.vstack(rows)
                info, rows = [], []
    if rows:
        yield info, np.vstack(rows)


def _gt_fast(body):
    """Biallelic diploid GT-only records whose calls are all 3 characters (`0/1`, `1|1`, `./.`): the sample part of the line
    is a regular n x 4 byte grid, decoded with numpy instead of a Python loop over samples.  None when the record is not regular."""
    if (len(body) + 1) % 4:
        return None
    a = np.frombuffer((body + "\t").encode("ascii", "replace"), dtype=np.uint8).reshape(-1, 4)
    x, y = a[:, 0], a[:, 2]
    if not (np.all((a[:, 1] == 47) | (a[:, 1] == 124)) and np.all(a[:, 3] == 9)
            and np.all(((x == 48) | (x == 49) | (x == 46)) & ((y == 48) | (y == 49) | (y == 46)))):
        return None
    return np.where((x == 46) | (y == 46), -1.0, (x == 49).astype(np.float64) + (y == 49))


def _ds_fast(body):
    """DS-only records without missing entries: the sample part parsed by numpy's C text reader.  None otherwise (a `.`
    entry, or anything the reader does not consume to the end, sends the record to the per-sample path)."""
    if "\t.\t" in body or body.startswith(".\t") or body.endswith("\t.") or body == ".":
        return None
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a = np.fromstring(body, dtype=np.float64, sep="\t")
    except ValueError:
        return None
    return a if a.size == body.count("\t") + 1 else None


def read_sample_file(path):
    """One ID per line without header (SPAGMMATtest's sampleFile), or an Oxford .sample file (two header lines, ID_2)."""
    lines = [l.split() for l in open(path) if l.strip()]
    if lines and lines[0][0] == "ID_1":
        return [l[1] if len(l) > 1 else l[0] for l in lines[2:]]
    return [l[0] for l in lines]


class BgenFile:
    def __init__(self, path):
        self.path = path
        self.f = open(path, "rb")
        offset, hlen, self.M, self.N = struct.unpack("<IIII", self.f.read(16))
        if self.f.read(4) not in (b"bgen", b"\0\0\0\0"):
            raise ValueError("%s is not a BGEN file" % path)
        self.f.seek(4 + hlen - 4)
        flags, = struct.unpack("<I", self.f.read(4))
        self.compression, self.layout, has_ids = flags & 3, (flags >> 2) & 15, flags >> 31
        if self.layout != 2:
            raise NotImplementedError("%s: BGEN layout %d (only v1.2, layout 2, is read; BGEN.cpp has the same limit)" % (path, self.layout))
        if self.compression not in (0, 1):
            raise NotImplementedError("%s: zstd-compressed BGEN is not read" % path)
        self.samples = None
        if has_ids:
            self.f.seek(4 + hlen)
            _, n = struct.unpack("<II", self.f.read(8))
            self.samples = []
            for _ in range(n):
                l, = struct.unpack("<H", self.f.read(2))
                self.samples.append(self.f.read(l).decode())
        self.f.seek(4 + offset)

    def _str(self, nbytes):
        l, = struct.unpack("<H" if nbytes == 2 else "<I", self.f.read(nbytes))
        return self.f.read(l).decode()

    def variants(self, allele_order="ref-first", chunk=1000, only=None, info_for=None):
        """Yields (info, D) like iter_vcf; D = copies of the tested allele (second allele for ref-first, first for alt-first).
        only: a set of "chr:pos:ref:alt" keys -- the blocks of every other variant are skipped without being inflated.
        info_for: boolean mask of samples; the imputation INFO scores (BGEN.cpp:275-345) of a chunk are then in self.last_info."""
        scores = []
        if allele_order not in ("ref-first", "alt-first"):
            raise ValueError("AlleleOrder should be 'ref-first' or 'alt-first'")
        info, rows = [], []
        for _ in range(self.M):
            vid = self._str(2)                # variant identifier
            rsid, chrom = self._str(2), self._str(2)
            if rsid == ".":                   # BGEN.cpp: RSID = (rsID == ".") ? snpID : rsID
                rsid = vid
            pos, K = struct.unpack("<IH", self.f.read(6))
            alleles = [self._str(4) for _ in range(K)]
            C, = struct.unpack("<I", self.f.read(4))
            if only is not None and K == 2:
                ref, alt = (alleles[1], alleles[0]) if allele_order == "alt-first" else (alleles[0], alleles[1])
                if "%s:%d:%s:%s" % (chrom, pos, ref, alt) not in only:
                    self.f.seek(C, 1)
                    continue
            if self.compression:
                D, = struct.unpack("<I", self.f.read(4))
                blk = zlib.decompress(self.f.read(C - 4))
                if len(blk) != D:
                    raise ValueError("%s: variant %s inflates to %d bytes, header says %d" % (self.path, rsid, len(blk), D))
            else:
                blk = self.f.read(C)
            if K != 2:
                raise NotImplementedError("%s: variant %s has %d alleles" % (self.path, rsid, K))
            n, k, pmin, pmax = struct.unpack("<IHBB", blk[:8])
            if n != self.N or k != 2:
                raise ValueError("%s: variant %s block is inconsistent with the header" % (self.path, rsid))
            pm = np.frombuffer(blk, dtype=np.uint8, count=n, offset=8)
            phased, bits = blk[8 + n], blk[9 + n]
            missing = pm >= 128
            if phased or np.any((pm & 63)[~missing] != 2):
                raise NotImplementedError("%s: variant %s is phased or not diploid (BGEN.cpp:172-177 stops on these too)" % (self.path, rsid))
            if bits == 8:
                p = np.frombuffer(blk, dtype=np.uint8, count=2 * n, offset=10 + n).astype(np.float64) / 255.0
            elif bits == 16:
                p = np.frombuffer(blk, dtype="<u2", count=2 * n, offset=10 + n).astype(np.float64) / 65535.0
            else:
                raise NotImplementedError("%s: %d-bit probabilities (8 and 16 are read; the reference reads 8 only)" % (self.path, bits))
            first = 2.0 * p[0::2] + p[1::2]                 # copies of the first allele (BGEN.cpp:218-223)
            if info_for is not None:
                use = np.asarray(info_for, dtype=bool) & ~missing
                cnt = float(use.sum())
                theta = first[use].sum() / (2 * cnt) if cnt else 0.0
                ff = (4.0 * p[0::2] + p[1::2])[use] - first[use] ** 2
                scores.append(1.0 if theta in (0.0, 1.0) else 1.0 - ff.sum() / (2 * cnt * theta * (1 - theta)))
            d = first if allele_order == "alt-first" else 2.0 - first
            d = np.where(missing, -1.0, d)
            ref, alt = (alleles[1], alleles[0]) if allele_order == "alt-first" else (alleles[0], alleles[1])
            info.append((chrom, str(pos), rsid, ref, alt))
            rows.append(d)
            if len(rows) == chunk:
                self.last_info = np.array(scores)
                yield info, np.vstack(rows)
                info, rows, scores = [], [], []
        if rows:
            self.last_info = np.array(scores)
            yield info, np.vstack(rows)

    def close(self):
        self.f.close()


class BgenNative:
    """The scan path: the library's host-side reader (csrc/bgen_reader.cu: sequential block reads, multi-threaded inflate +
    decode).  Same interface as BgenFile, which stays as the independent pure-Python implementation the tests compare it with."""

    def __init__(self, path, n_threads=0):
        import ctypes as C
        import os
        from . import _lib
        self._L, self._C = _lib.lib(), C
        h, n, m, ids = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int()
        if self._L.sgb_bgen_open(path.encode(), C.byref(h), C.byref(n), C.byref(m), C.byref(ids)):
            raise ValueError(self._L.sgb_last_error(None).decode())
        self._h, self.N, self.M, self.path = h, n.value, m.value, path
        self.n_threads = n_threads or min(16, os.cpu_count() or 1)
        self.samples = None
        if ids.value:
            buf = C.create_string_buffer(1 << 16)
            self.samples = []
            for i in range(self.N):
                if self._L.sgb_bgen_sample_id(h, i, buf, len(buf)):
                    raise ValueError(self._L.sgb_last_error(None).decode())
                self.samples.append(buf.value.decode())

    def variants(self, allele_order="ref-first", chunk=1000, info_for=None):
        """info_for: boolean mask of samples; the INFO scores of the chunk just yielded are then in self.last_info."""
        if allele_order not in ("ref-first", "alt-first"):
            raise ValueError("AlleleOrder should be 'ref-first' or 'alt-first'")
        C = self._C
        info_buf = C.create_string_buffer(chunk * 1024)
        mask = None if info_for is None else np.ascontiguousarray(info_for, dtype=np.uint8)
        while True:
            D = np.empty((chunk, self.N))
            scores = np.empty(chunk) if mask is not None else None
            got = C.c_int64()
            if self._L.sgb_bgen_read(self._h, chunk, int(allele_order == "alt-first"), self.n_threads,
                                     None if mask is None else mask.ctypes.data, D.ctypes.data,
                                     None if scores is None else scores.ctypes.data, info_buf, len(info_buf), C.byref(got)):
                raise ValueError(self._L.sgb_last_error(None).decode())
            if got.value == 0:
                return
            info = [tuple(l.split("\t")) for l in info_buf.value.decode().split("\n") if l]
            self.last_info = None if scores is None else scores[:got.value]
            yield info, D[:got.value]

    def close(self):
        if self._h:
            self._L.sgb_bgen_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hardcalls_to_bed_rows(D):
    """Integral dosage rows -> raw PLINK rows (2 bits per sample, A1 = the tested allele: 00 = 2 copies, 10 = 1, 11 = 0,
    01 = missing; PLINK.hpp:48-56).  Raises when a dosage is not 0 / 1 / 2 / missing: fractional dosages need the dosage entry."""
    D = np.asarray(D)
    g = np.where(D < 0, -1, np.rint(D)).astype(np.int64)
    if np.any((D >= 0) & (np.abs(D - g) > 1e-12)) or g.max(initial=0) > 2:
        raise ValueError("fractional dosages cannot be packed as hard calls")
    code = np.array([3, 2, 0, 1], dtype=np.uint8)[g]          # 0 -> 11, 1 -> 10, 2 -> 00, -1 -> 01
    nm, n = code.shape
    B0 = (n + 3) // 4
    pad = np.full((nm, B0 * 4), 3, dtype=np.uint8)
    pad[:, :n] = code
    q = pad.reshape(nm, B0, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8).reshape(-1)
