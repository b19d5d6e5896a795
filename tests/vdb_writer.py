"""TEST INFRASTRUCTURE: a minimal OpenVDB file WRITER (pure Python) so that the .vdb reader
in libcvr_b200.so can be exercised on boxes where the reference checkout (and its
data/vdb/bonsai_small.vdb) is not mounted.

It emits the same on-disk structures the reader decodes -- file version 224, 5-4-3 trees of
float / vec3s, per-grid compression flags, node-mask compression codes 0..6, and chunks that
are stored raw, zlib-deflated or wrapped in a c-blosc 1.x container with byte shuffle, the
block-split rule and LZ4 block streams (a small greedy LZ4 compressor is included so that
match copies, overlapping matches and long literal/match runs occur).  The reader itself is
pinned against the real bonsai_small.vdb in this container (tests/test_vdb.py,
tests/golden/vdb_bonsai_small.json)."""
from __future__ import annotations

import struct
import zlib

import numpy as np

COMPRESS_ZIP, COMPRESS_ACTIVE_MASK, COMPRESS_BLOSC = 1, 2, 4


# ---------------------------------------------------------------- LZ4 block compressor (greedy)
def lz4_compress(src: bytes) -> bytes:
    n = len(src)
    out = bytearray()
    table: dict[bytes, int] = {}
    anchor = 0
    i = 0

    def emit(lit: bytes, mlen: int | None, off: int = 0):
        ll = len(lit)
        tok_l = min(ll, 15)
        tok_m = 0 if mlen is None else min(mlen - 4, 15)
        out.append((tok_l << 4) | tok_m)
        if ll >= 15:
            r = ll - 15
            while r >= 255:
                out.append(255)
                r -= 255
            out.append(r)
        out.extend(lit)
        if mlen is not None:
            out.extend(struct.pack("<H", off))
            if mlen - 4 >= 15:
                r = mlen - 4 - 15
                while r >= 255:
                    out.append(255)
                    r -= 255
                out.append(r)

    # LZ4 end conditions: the last 5 bytes are literals, the last match starts >= 12 bytes before the end
    limit = n - 12
    while i < limit:
        key = src[i:i + 4]
        cand = table.get(key)
        table[key] = i
        if cand is not None and i - cand <= 65535:
            m = 4
            while i + m < n - 5 and src[cand + m] == src[i + m]:
                m += 1
            emit(src[anchor:i], m, i - cand)
            i += m
            anchor = i
        else:
            i += 1
    emit(src[anchor:], None)
    return bytes(out)


# ---------------------------------------------------------------- c-blosc 1.x container
def blosc_compress(data: bytes, typesize: int, blocksize: int | None = None, shuffle: bool = True) -> bytes | None:
    """Returns a blosc chunk, or None when it would not be smaller (OpenVDB then stores raw)."""
    nbytes = len(data)
    if nbytes == 0:
        return None
    if blocksize is None:
        blocksize = nbytes
    flags = (1 if shuffle else 0) | (1 << 5)  # byte shuffle + LZ4
    nblocks = (nbytes + blocksize - 1) // blocksize
    body = bytearray()
    starts = []
    base = 16 + 4 * nblocks
    for b in range(nblocks):
        blk = data[b * blocksize:(b + 1) * blocksize]
        bsize = len(blk)
        leftover = bsize != blocksize
        if shuffle and typesize > 1:
            ne = bsize // typesize
            a = np.frombuffer(blk[:ne * typesize], np.uint8).reshape(ne, typesize).T.copy().tobytes()
            blk = a + blk[ne * typesize:]
        split = typesize <= 16 and (blocksize // typesize) >= 128 and not leftover
        nsplits = typesize if split else 1
        neblock = bsize // nsplits
        starts.append(base + len(body))
        for j in range(nsplits):
            s = blk[j * neblock:(j + 1) * neblock] if nsplits > 1 else blk
            c = lz4_compress(s)
            if len(c) >= len(s):
                c = s  # stored raw: compressed size == plain size
            body.extend(struct.pack("<i", len(c)))
            body.extend(c)
    chunk = bytearray(struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, base + len(body)))
    for s in starts:
        chunk.extend(struct.pack("<i", s))
    chunk.extend(body)
    return bytes(chunk) if len(chunk) < nbytes else None


def blosc_stored(data: bytes, typesize: int) -> bytes:
    """The 'memcpyed' form of a blosc chunk (flag 0x2)."""
    return struct.pack("<BBBBIII", 2, 1, 0x2 | 1, typesize, len(data), len(data), len(data) + 16) + data


# ---------------------------------------------------------------- OpenVDB pieces
def _s(x: str) -> bytes:
    b = x.encode()
    return struct.pack("<I", len(b)) + b


def _meta(entries) -> bytes:
    out = struct.pack("<I", len(entries))
    for name, typ, payload in entries:
        out += _s(name) + _s(typ) + struct.pack("<I", len(payload)) + payload
    return out


class _GridWriter:
    def __init__(self, name, channels, background, leaves, compression, blosc_blocksize=None, stored_every=0,
                 half=False):
        self.name, self.C, self.bg = name, channels, np.asarray(background, np.float32).reshape(channels)
        self.leaves = leaves  # list of (origin(3), mask(512 bool), values(512, C) float32)
        self.comp = compression
        self.blocksize = blosc_blocksize
        self.stored_every = stored_every
        self.half = half
        self._chunk_no = 0

    # io::writeData
    def data(self, vals: np.ndarray) -> bytes:
        raw = (vals.astype(np.float16) if self.half else vals.astype(np.float32)).tobytes()
        elem = (2 if self.half else 4) * self.C
        if self.comp & COMPRESS_BLOSC:
            self._chunk_no += 1
            c = None
            if self.stored_every and self._chunk_no % self.stored_every == 0 and raw:
                c = blosc_stored(raw, elem)
            elif raw:
                c = blosc_compress(raw, elem, self.blocksize)
            if c is None:
                return struct.pack("<q", -len(raw)) + raw
            return struct.pack("<q", len(c)) + c
        if self.comp & COMPRESS_ZIP:
            c = zlib.compress(raw) if raw else b""
            if not raw or len(c) >= len(raw):
                return struct.pack("<q", -len(raw)) + raw
            return struct.pack("<q", len(c)) + c
        return raw

    # io::writeCompressedValues: picks the node-mask compression code like OpenVDB does
    def compressed(self, vals: np.ndarray, mask: np.ndarray) -> bytes:
        vals = vals.reshape(-1, self.C)
        n = len(mask)
        if not (self.comp & COMPRESS_ACTIVE_MASK):
            return struct.pack("<b", 6) + self.data(vals)
        ina = vals[~mask]
        uniq = [tuple(u) for u in np.unique(ina, axis=0)] if len(ina) else []
        bg, nbg = tuple(self.bg), tuple(-self.bg)
        sel = None
        extra = b""
        if not uniq or uniq == [bg]:
            code = 0
        elif uniq == [nbg]:
            code = 1
        elif len(uniq) == 1:
            code, extra = 2, np.asarray(uniq[0], np.float32).tobytes()
        elif len(uniq) == 2 and set(uniq) == {bg, nbg}:
            code = 3
            sel = np.all(vals == self.bg, axis=1) & ~mask  # ON -> +background
        elif len(uniq) == 2 and bg in uniq:
            other = uniq[0] if uniq[1] == bg else uniq[1]
            code, extra = 4, np.asarray(other, np.float32).tobytes()
            sel = np.all(vals == self.bg, axis=1) & ~mask  # ON -> background, OFF -> the other value
        elif len(uniq) == 2:
            code = 5
            extra = np.asarray(uniq[0], np.float32).tobytes() + np.asarray(uniq[1], np.float32).tobytes()
            sel = np.all(vals == np.asarray(uniq[1], np.float32), axis=1) & ~mask  # ON -> second value
        else:
            code = 6
        out = struct.pack("<b", code) + extra
        if sel is not None:
            out += _bits(sel)
        if code == 6:
            return out + self.data(vals)
        return out + self.data(vals[mask])

    def topology_and_buffers(self) -> tuple[bytes, bytes]:
        C = self.C
        roots: dict[tuple, dict[int, dict[int, tuple]]] = {}
        for org, mask, vals in self.leaves:
            org = np.asarray(org, np.int64)
            r = tuple((org // 4096) * 4096)
            rel = org - np.asarray(r)
            i5 = int(((rel[0] // 128) << 10) | ((rel[1] // 128) << 5) | (rel[2] // 128))
            rel4 = rel % 128
            i4 = int(((rel4[0] // 8) << 8) | ((rel4[1] // 8) << 4) | (rel4[2] // 8))
            roots.setdefault(r, {}).setdefault(i5, {})[i4] = (mask, vals)
        top = struct.pack("<i", 1) + self.bg.astype(np.float32).tobytes() + struct.pack("<II", 0, len(roots))
        buf = b""
        for r in sorted(roots):  # std::map<Coord,...> order: lexicographic x, y, z
            top += struct.pack("<iii", *r)
            n5 = roots[r]
            child = np.zeros(32768, bool)
            child[list(n5)] = True
            top += _bits(child) + _bits(np.zeros(32768, bool))
            top += self.compressed(np.tile(self.bg, (32768, 1)), np.zeros(32768, bool))
            for i5 in sorted(n5):
                n4 = n5[i5]
                child4 = np.zeros(4096, bool)
                child4[list(n4)] = True
                top += _bits(child4) + _bits(np.zeros(4096, bool))
                top += self.compressed(np.tile(self.bg, (4096, 1)), np.zeros(4096, bool))
                for i4 in sorted(n4):
                    mask, vals = n4[i4]
                    top += _bits(mask)
                    buf += _bits(mask) + self.compressed(np.asarray(vals, np.float32).reshape(512, C), mask)
        return top, buf


def _bits(mask: np.ndarray) -> bytes:
    """util::NodeMask words: bit n of the mask -> bit (n & 63) of 64-bit word n >> 6."""
    return np.packbits(mask.astype(np.uint8), bitorder="little").tobytes()


def leaves_from_dense(dense: np.ndarray, active: np.ndarray, origin=(0, 0, 0), inactive_fill=None):
    """dense: (nz, ny, nx[, C]) -> list of leaves over 8^3 blocks that contain an active voxel.
    Index-space coordinate of dense[z, y, x] is origin + (x, y, z).  inactive_fill(leaf_no, mask)
    may return (512, C) values to use for the inactive voxels of a leaf."""
    if dense.ndim == 3:
        dense = dense[..., None]
    nz, ny, nx, C = dense.shape
    ox, oy, oz = origin
    out = []
    x0, y0, z0 = (ox // 8) * 8, (oy // 8) * 8, (oz // 8) * 8
    k = 0
    for bx in range(x0, ox + nx, 8):
        for by in range(y0, oy + ny, 8):
            for bz in range(z0, oz + nz, 8):
                vals = np.zeros((8, 8, 8, C), np.float32)  # [x, y, z]
                mask = np.zeros((8, 8, 8), bool)
                xs, ys, zs = max(bx, ox), max(by, oy), max(bz, oz)
                xe, ye, ze = min(bx + 8, ox + nx), min(by + 8, oy + ny), min(bz + 8, oz + nz)
                sub = dense[zs - oz:ze - oz, ys - oy:ye - oy, xs - ox:xe - ox]  # [z, y, x]
                sa = active[zs - oz:ze - oz, ys - oy:ye - oy, xs - ox:xe - ox]
                vals[xs - bx:xe - bx, ys - by:ye - by, zs - bz:ze - bz] = sub.transpose(2, 1, 0, 3)
                mask[xs - bx:xe - bx, ys - by:ye - by, zs - bz:ze - bz] = sa.transpose(2, 1, 0)
                if not mask.any():
                    continue
                m = mask.reshape(512)
                v = vals.reshape(512, C)
                v[~m] = 0.0
                if inactive_fill is not None:
                    f = inactive_fill(k, m)
                    if f is not None:
                        v[~m] = np.asarray(f, np.float32).reshape(512, C)[~m]
                out.append(((bx, by, bz), m, v))
                k += 1
    return out


def write_vdb(path: str, grids, compression: int = COMPRESS_BLOSC | COMPRESS_ACTIVE_MASK, blosc_blocksize=None,
              stored_every: int = 0, extra_grid: bool = False, half: bool = False) -> None:
    """grids: list of (name, channels, background, leaves)."""
    head = struct.pack("<q", 0x56444220) + struct.pack("<III", 224, 8, 1) + b"\x01" + b"0" * 8 + b"-" + b"0" * 4 + b"-" + \
        b"0" * 4 + b"-" + b"0" * 4 + b"-" + b"0" * 12
    head += _meta([("creator", "string", b"tests/vdb_writer.py")])
    specs = list(grids)
    if extra_grid:  # a grid type the reader must skip by offset
        specs = [("ids", None, None, None)] + specs
    head += struct.pack("<I", len(specs))
    body = b""
    pos = len(head)
    for name, C, bg, leaves in specs:
        if C is None:
            desc = _s(name) + _s("Tree_int32_5_4_3") + _s("")
            payload = b"\xAB" * 37
            gpos = pos + len(desc) + 24
            body += desc + struct.pack("<qqq", gpos, gpos + 5, gpos + len(payload)) + payload
            pos = gpos + len(payload)
            continue
        typ = "Tree_float_5_4_3" if C == 1 else "Tree_vec3s_5_4_3"
        desc = _s(name) + _s(typ + ("_HalfFloat" if half else "")) + _s("")
        gw = _GridWriter(name, C, bg, leaves, compression, blosc_blocksize, stored_every, half)
        top, buf = gw.topology_and_buffers()
        n_active = int(sum(int(m.sum()) for _, m, _ in leaves))
        pre = struct.pack("<I", compression)
        pre += _meta([("class", "string", b"fog volume"), ("name", "string", name.encode()),
                      ("file_voxel_count", "int64", struct.pack("<q", n_active)),
                      ("is_saved_as_half_float", "bool", b"\x01" if half else b"\x00"),
                      ("__delayedload", "__delayedload", b"\x00" * 23)])
        pre += _s("UniformScaleMap") + struct.pack("<15d", *([0.01] * 3 + [0.01] * 3 + [100.0] * 3 + [1e4] * 3 + [50.0] * 3))
        gpos = pos + len(desc) + 24
        bpos = gpos + len(pre) + len(top)
        epos = bpos + len(buf)
        body += desc + struct.pack("<qqq", gpos, bpos, epos) + pre + top + buf
        pos = epos
    with open(path, "wb") as f:
        f.write(head + body)
