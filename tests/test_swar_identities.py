"""CPU: the integer identities behind the packed-lane device code (docs/ARITHMETIC.md,
"implementation notes"), restated with numpy on 32-bit words and checked against the plain
definitions -- exhaustively where the domain is small, on random words otherwise.

  * frame kernel, 2x2 reduction on 16-bit lanes (uwt_frame_fused.cu: pair_sums / down4)
  * threshold test of 4 pixels (g > thr per byte) and the 8-row column masks of the scatter
  * depth tests of 4 pixels from one vector load (uwt_internal.cuh: depth_nz4, valid4)
  * packed record fields (pack_record / the scatter's gx | gy << 13)
"""
import numpy as np

U32 = np.uint32
M32 = 0xFFFFFFFF


def byte_perm(x, y, sel):
    """PTX prmt.b32 (default mode) on arrays of uint32."""
    x, y = np.asarray(x, np.uint64), np.asarray(y, np.uint64)
    both = x | (y << np.uint64(32))
    out = np.zeros_like(x)
    for i in range(4):
        k = (sel >> (4 * i)) & 7
        out |= ((both >> np.uint64(8 * k)) & np.uint64(0xFF)) << np.uint64(8 * i)
    return out.astype(U32)


def bytes_of(w):
    w = np.asarray(w, U32)
    return np.stack([(w >> U32(8 * i)) & U32(0xFF) for i in range(4)], -1).astype(np.int64)


def test_pair_sums_and_down4():
    rng = np.random.default_rng(1)
    n = 200_000
    ax, ay, bx, by = (rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(U32) for _ in range(4))
    ax[:4] = ay[:4] = bx[:4] = by[:4] = [0, M32, 0x00FF00FF, 0xFF00FF00]

    def pair_sums(x):
        return (byte_perm(x, 0, 0x4240).astype(np.uint64) +
                byte_perm(x, 0, 0x4341).astype(np.uint64)).astype(U32)
    s0 = (pair_sums(ax).astype(np.uint64) + pair_sums(bx) + 0x00020002).astype(U32)
    s1 = (pair_sums(ay).astype(np.uint64) + pair_sums(by) + 0x00020002).astype(U32)
    out = byte_perm(s0 >> U32(2), s1 >> U32(2), 0x6420)
    a = np.concatenate([bytes_of(ax), bytes_of(ay)], -1)     # 8 pixels of the upper row
    b = np.concatenate([bytes_of(bx), bytes_of(by)], -1)
    ref = (a[:, 0::2] + a[:, 1::2] + b[:, 0::2] + b[:, 1::2] + 2) >> 2   # System.cpp:246-251
    assert np.array_equal(bytes_of(out), ref)


def vsetgtu4(a, b):
    return sum(((bytes_of(a)[..., i] > bytes_of(b)[..., i]).astype(np.uint64) << np.uint64(8 * i))
               for i in range(4)).astype(U32)


def test_threshold_bytes_and_column_masks():
    rng = np.random.default_rng(2)
    g = rng.integers(0, 256, (5000, 8, 4)).astype(np.int64)           # [group, row, pixel]
    for thr in (0, 19, 20, 77, 254, 255, 300):
        t = min(thr, 255)
        acc = np.zeros(g.shape[0], np.uint64)
        for row in range(8):
            word = sum(g[:, row, i].astype(np.uint64) << np.uint64(8 * i) for i in range(4))
            sel4 = vsetgtu4(word.astype(U32), np.full(g.shape[0], t * 0x01010101, U32))
            acc |= sel4.astype(np.uint64) << np.uint64(row)        # acc |= sel4 << row
        acc = acc.astype(U32)
        for i in range(4):                                         # byte i = mask of column i
            mask = bytes_of(acc)[:, i]
            ref = sum(((g[:, row, i] > thr).astype(np.int64) << row) for row in range(8))
            assert np.array_equal(mask, ref), (thr, i)
    # a record's slot inside its run: popc(mask & below(row)) = selected rows above it
    m = rng.integers(0, 256, 1000)
    for rq in range(8):
        below = (1 << rq) - 1
        assert np.array_equal([bin(int(v) & below).count("1") for v in m],
                              [sum((int(v) >> r) & 1 for r in range(rq)) for v in m])
    # nibble gather of the level-0 bitmask: (m * 0x01020408) >> 24 for m with 0 / 1 bytes
    for bits in range(16):
        word = sum(((bits >> i) & 1) << (8 * i) for i in range(4))
        assert ((word * 0x01020408) & M32) >> 24 == bits


def test_depth_nz4_and_valid4():
    rng = np.random.default_rng(3)
    n = 100_000
    d = rng.integers(0, 1 << 16, (n, 4)).astype(np.int64)
    d[rng.random((n, 4)) < 0.3] = 0
    d[:8] = [[0, 0, 0, 0], [1, 0, 0x8000, 0xFFFF], [0x00FF, 0xFF00, 0x0100, 0x0001],
             [0x7FFF, 0x8001, 0, 0x8000], [0, 1, 0, 1], [0x8000] * 4, [0xFFFF] * 4, [0x0080] * 4]
    lo = (d[:, 0] | (d[:, 1] << 16)).astype(U32)
    hi = (d[:, 2] | (d[:, 3] << 16)).astype(U32)

    def nz2(w, positive):
        w64 = w.astype(np.uint64)
        t = ((w64 | ((w64 & 0x7FFF7FFF) + 0x7FFF7FFF)) & M32) >> np.uint64(15)
        if positive:
            t &= ~(w64 >> np.uint64(15)) & np.uint64(M32)
        t &= np.uint64(0x00010001)
        return ((t | (t >> np.uint64(8))) & np.uint64(0x0101)).astype(U32)
    for positive in (False, True):
        out = nz2(lo, positive) | (nz2(hi, positive) << U32(16))
        signed = np.where(d >= 0x8000, d - 0x10000, d)
        ref = (signed > 0) if positive else (d != 0)              # Tracker.cpp:1273 / :1339
        assert np.array_equal(bytes_of(out), ref.astype(np.int64)), positive
    # REFERENCE mode: at<uchar>(y, x) on the 16-bit row = byte x (Tracker.cpp:1339)
    v = rng.integers(0, 1 << 32, n, dtype=np.uint64)
    v[rng.random(n) < 0.3] &= 0x00FF00FF
    out = (((v | ((v & 0x7F7F7F7F) + 0x7F7F7F7F)) >> np.uint64(7)) & np.uint64(0x01010101)).astype(U32)
    assert np.array_equal(bytes_of(out), (bytes_of(v.astype(U32)) != 0).astype(np.int64))
    # valid4: 0x01 in the bytes of the pixels gx .. gx + 3 inside a row of width w
    for w in range(0, 12):
        for gx in range(0, 12, 4):
            k = w - gx
            got = 0x01010101 if k >= 4 else (0 if k <= 0 else 0x01010101 >> (8 * (4 - k)))
            assert [(got >> (8 * i)) & 0xFF for i in range(4)] == \
                [1 if gx + i < w else 0 for i in range(4)]


def test_record_packing_fields():
    """pack_record (uwt_internal.cuh) and the scatter's two halves: x:12 | y:12 | I1:8 and
    gx:13 | gy:13 (two's complement, |g| <= 16 * 255), unpacked the way the sweep does."""
    rng = np.random.default_rng(4)
    n = 100_000
    x, y = rng.integers(0, 4096, n), rng.integers(0, 4096, n)
    i1 = rng.integers(0, 256, n)
    gx, gy = rng.integers(-4080, 4081, n), rng.integers(-4080, 4081, n)
    lo = (x | (y << 12) | (i1 << 24)).astype(np.uint64)
    gx13 = (gx & 0x1FFF).astype(np.uint64)
    gy13 = (gy & 0x1FFF).astype(np.uint64)
    hi = gx13 | (gy13 << np.uint64(13))
    rec = lo | (hi << np.uint64(32))
    ref = ((x & 0xFFF) | ((y & 0xFFF) << 12) | ((i1 & 0xFF) << 24)).astype(np.uint64) | \
        (gx13 << np.uint64(32)) | (gy13 << np.uint64(45))
    assert np.array_equal(rec, ref)
    h32 = (rec >> np.uint64(32)).astype(np.int64)
    sx = ((h32 << 19) & M32).astype(np.uint32).view(np.int32) >> 19      # ((int)(hi << 19)) >> 19
    sy = ((h32 << 6) & M32).astype(np.uint32).view(np.int32) >> 19       # ((int)(hi << 6)) >> 19
    assert np.array_equal(sx, gx) and np.array_equal(sy, gy)
    l32 = (rec & np.uint64(M32)).astype(np.int64)
    assert np.array_equal(l32 & 0xFFF, x) and np.array_equal((l32 >> 12) & 0xFFF, y)
    assert np.array_equal(l32 >> 24, i1)
