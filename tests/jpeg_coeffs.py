"""Minimal baseline-JPEG parser (test helper): quantised DCT coefficients, quant tables and a float
reconstruction of a Huffman-coded sequential JPEG. Used to pin the oracle's VarDCT path on the
reference's sample.jpg <-> sample_jpg.jxl pair (golden G3)."""
import numpy as np

ZIGZAG = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
          28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
          54, 47, 55, 62, 63]


class _Bits:
    def __init__(self, data):
        self.d, self.pos, self.acc, self.n = data, 0, 0, 0

    def bit(self):
        if self.n == 0:
            b = self.d[self.pos]
            self.pos += 1
            if b == 0xFF:
                assert self.d[self.pos] == 0, "unexpected marker in scan"
                self.pos += 1
            self.acc, self.n = b, 8
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, k):
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v


def _huff_table(counts, symbols):
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


def _decode_symbol(br, table):
    code = 0
    for length in range(1, 17):
        code = (code << 1) | br.bit()
        if (length, code) in table:
            return table[(length, code)]
    raise ValueError("bad huffman code")


def _extend(v, t):
    return v if v >= (1 << (t - 1)) else v - (1 << t) + 1


def parse(data: bytes):
    """Returns dict(width, height, comps=[{id,h,v,tq,coeffs[by,bx,64 natural order]}], qt={id: 64 natural order})."""
    assert data[:2] == b"\xff\xd8"
    pos = 2
    qt, dc_tabs, ac_tabs = {}, {}, {}
    frame = None
    while pos < len(data):
        assert data[pos] == 0xFF
        marker = data[pos + 1]
        pos += 2
        if marker == 0xD9:
            break
        length = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2:pos + length]
        pos += length
        if marker == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                assert pq == 0
                tbl = np.zeros(64, np.int32)
                for k in range(64):
                    tbl[ZIGZAG[k]] = seg[i + 1 + k]
                qt[tq] = tbl
                i += 65
        elif marker == 0xC0:
            assert seg[0] == 8
            h, w, n = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [dict(id=seg[6 + 3 * i], h=seg[7 + 3 * i] >> 4, v=seg[7 + 3 * i] & 15, tq=seg[8 + 3 * i])
                     for i in range(n)]
            frame = dict(width=w, height=h, comps=comps)
        elif marker in (0xC1, 0xC2, 0xC3):
            raise ValueError("only baseline JPEG is supported by this helper")
        elif marker == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1:i + 17])
                n = sum(counts)
                symbols = list(seg[i + 17:i + 17 + n])
                (dc_tabs if tc == 0 else ac_tabs)[th] = _huff_table(counts, symbols)
                i += 17 + n
        elif marker == 0xDA:
            ns = seg[0]
            sel = {seg[1 + 2 * i]: (seg[2 + 2 * i] >> 4, seg[2 + 2 * i] & 15) for i in range(ns)}
            comps = frame["comps"]
            hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
            mcux = -(-frame["width"] // (8 * hmax))
            mcuy = -(-frame["height"] // (8 * vmax))
            for c in comps:
                c["coeffs"] = np.zeros((mcuy * c["v"], mcux * c["h"], 64), np.int32)
            br = _Bits(data[pos:])
            pred = {c["id"]: 0 for c in comps}
            for my in range(mcuy):
                for mx in range(mcux):
                    for c in comps:
                        td, ta = sel[c["id"]]
                        for v in range(c["v"]):
                            for h in range(c["h"]):
                                blk = c["coeffs"][my * c["v"] + v, mx * c["h"] + h]
                                t = _decode_symbol(br, dc_tabs[td])
                                diff = _extend(br.bits(t), t) if t else 0
                                pred[c["id"]] += diff
                                blk[0] = pred[c["id"]]
                                k = 1
                                while k < 64:
                                    rs = _decode_symbol(br, ac_tabs[ta])
                                    r, s = rs >> 4, rs & 15
                                    if s == 0:
                                        if r != 15:
                                            break
                                        k += 16
                                        continue
                                    k += r
                                    blk[ZIGZAG[k]] = _extend(br.bits(s), s)
                                    k += 1
            frame["qt"] = qt
            return frame
    raise ValueError("no scan found")


def reconstruct_rgb_float(frame):
    """Float reconstruction (exact IDCT, JFIF YCbCr -> RGB), no rounding or clamping; 4:4:4 only."""
    comps = frame["comps"]
    assert all(c["h"] == 1 and c["v"] == 1 for c in comps)
    n = np.arange(8)
    basis = np.cos(np.pi * n[None, :] * (n[:, None] + 0.5) / 8) * np.where(n[None, :] > 0, 1.0, np.sqrt(0.5)) * 0.5
    planes = []
    for c in comps:
        by, bx, _ = c["coeffs"].shape
        out = np.zeros((by * 8, bx * 8))
        q = frame["qt"][c["tq"]].astype(np.float64)
        for y in range(by):
            for x in range(bx):
                f = (c["coeffs"][y, x] * q).reshape(8, 8)  # [v][u]
                out[y * 8:y * 8 + 8, x * 8:x * 8 + 8] = basis @ f @ basis.T
        planes.append(out + (128.0 if len(planes) == 0 else 0.0))
    yy, cb, cr = planes
    r = yy + 1.402 * cr
    g = yy - (0.114 * 1.772 / 0.587) * cb - (0.299 * 1.402 / 0.587) * cr
    b = yy + 1.772 * cb
    h, w = frame["height"], frame["width"]
    return np.stack([r, g, b], axis=2)[:h, :w]
