"""Exhaustive proofs of the closed forms the CUDA kernel uses in place of the reference's
rounded SSE2 cascades (goofy_b200/csrc/block_codec.cuh header)."""
import numpy as np


def avg(a, b):
    return (a + b + 1) >> 1


def test_luma_closed_form_all_2_24():
    """avg(avg(R,B),G) == (R + 2G + B + 3) >> 2 for every (R,G,B)  (goofy_tc.h:1166)."""
    r = np.arange(256, dtype=np.int32).reshape(256, 1, 1)
    g = np.arange(256, dtype=np.int32).reshape(1, 256, 1)
    b = np.arange(256, dtype=np.int32).reshape(1, 1, 256)
    assert np.array_equal(avg(avg(r, b), g) + 0 * r, (r + 2 * g + b + 3) >> 2)


def test_to5_closed_form():
    """avg(avg(avg(sat-(v,8),0),0),0) == (max(v,1)-1) >> 3  (goofy_tc.h:1309-1311)."""
    v = np.arange(256, dtype=np.int32)
    cascade = avg(avg(avg(np.maximum(v - 8, 0), 0), 0), 0)
    assert np.array_equal(cascade, (np.maximum(v, 1) - 1) >> 3)
    assert cascade.max() == 31


def test_quant_threshold_closed_form():
    """quarter + eighth of the rounded halvings == ((r+3)>>2) + ((r+7)>>3), never saturates  (goofy_tc.h:1184-1190)."""
    r = np.arange(8, 256, dtype=np.int32)
    half = avg(r, 0); quarter = avg(half, 0); eighth = avg(quarter, 0)
    qt = np.minimum(quarter + eighth, 255)
    assert np.array_equal(qt, ((r + 3) >> 2) + ((r + 7) >> 3))
    assert qt.min() == 3 and qt.max() == 96


def test_ceil_avg_via_complemented_floor_avg():
    a = np.arange(256, dtype=np.int32).reshape(256, 1)
    b = np.arange(256, dtype=np.int32).reshape(1, 256)
    na, nb = 255 - a, 255 - b
    floor_avg = (na & nb) + (((na ^ nb) & 0xFE) >> 1)
    assert np.array_equal(255 - floor_avg, avg(a, b) + 0 * a)


def test_classification_in_unshifted_domain():
    """Gez / Lqt computed from S = R+2G+B+3 and e = S - 4*mid equal the reference's byte-lane form
    (goofy_tc.h:1212-1224) for every brightness, mid and threshold that can occur."""
    S = np.arange(3, 1024, dtype=np.int32).reshape(-1, 1, 1)     # S = R+2G+B+3
    mid = np.arange(0, 256, dtype=np.int32).reshape(1, -1, 1)
    qt = np.arange(3, 97, dtype=np.int32).reshape(1, 1, -1)
    y = S >> 2
    pos = np.minimum(np.maximum(y - mid, 0), 127)
    neg = np.minimum(np.maximum(mid - y, 0), 127)
    gez_ref = neg == 0
    lqt_ref = (pos | neg) < qt
    e = S - 4 * mid
    assert np.array_equal(gez_ref + 0 * qt, (e >= 0) + 0 * qt)
    assert np.array_equal(lqt_ref, (e >= 4 - 4 * qt) & (e < 4 * qt))
    # the packed-lane encoding: lane = e + 0x4000; bit 14 <=> gez; bit 15 of (lane+kLo) ^ (lane+kHi) <=> lqt
    lane = e + 0x4000
    assert lane.min() > 0 and lane.max() < 0x8000
    k_lo, k_hi = 0x3FFC + 4 * qt, 0x4000 - 4 * qt
    assert (lane + k_lo).max() < 0x10000 and (lane + k_hi).min() >= 0
    assert np.array_equal(((lane >> 14) & 1).astype(bool) + 0 * qt, gez_ref + 0 * qt)
    assert np.array_equal(((((lane + k_lo) ^ (lane + k_hi)) >> 15) & 1).astype(bool), lqt_ref)


def test_etc1_control_table_steps():
    """The table at goofy_tc.h:1040-1057 as thresholds; checked against the table itself when the reference is mounted."""
    steps = [22, 44, 74, 106, 152, 182, 254]
    table = [(sum(r >= s for s in steps) * 36 + 3) << 24 for r in range(256)]
    assert table[0] == 0x03000000 and table[21] == 0x03000000 and table[22] == 0x27000000
    assert table[253] == 0xDB000000 and table[254] == 0xFF000000 and table[255] == 0xFF000000
    import re
    from pathlib import Path
    hdr = Path("/root/reference/GoofyTC/goofy_tc.h")
    if hdr.exists():
        txt = hdr.read_text()
        body = txt[txt.index("etc1BrighnessRangeTocontrolByte[256]"):]
        body = body[: body.index("};")]
        vals = [int(v, 16) for v in re.findall(r"0x[0-9A-Fa-f]{8}", body)]
        assert vals == table


def test_base_colour_correction_as_signed_clamp():
    """up ? sat+(a,corr) : sat-(a,corr) with corr = min(|mid-aY|,127) == clamp(a + clamp(mid-aY,-127,127), 0, 255)."""
    a = np.arange(256, dtype=np.int32).reshape(-1, 1, 1)
    mid = np.arange(256, dtype=np.int32).reshape(1, -1, 1)
    ay = np.arange(256, dtype=np.int32).reshape(1, 1, -1)
    pos = np.minimum(np.maximum(mid - ay, 0), 127)
    neg = np.minimum(np.maximum(ay - mid, 0), 127)
    corr = pos | neg
    ref = np.where(neg == 0, np.minimum(a + corr, 255), np.maximum(a - corr, 0))
    d = np.clip(mid - ay, -127, 127)
    assert np.array_equal(ref, np.clip(a + d, 0, 255))


def test_flag_byte_lanes_and_weights():
    """The flag-byte selector scheme (block_codec.cuh, selectors_from_flag_bytes): the three threshold flags are the
    sign bits of u16 lanes that never exchange a carry, and the IDP weights rebuild the DXT1 index byte and the
    ETC1s plane bytes from them -- for every brightness, mid and threshold that can occur."""
    y4 = np.arange(0, 1021, dtype=np.int64).reshape(-1, 1, 1)      # R + 2G + B
    mid = np.arange(0, 256, dtype=np.int64).reshape(1, -1, 1)
    qt = np.arange(3, 97, dtype=np.int64).reshape(1, 1, -1)
    q4 = 4 * qt
    e = y4 + 3 - 4 * mid
    g_lane = y4 + (0x8003 - 4 * mid) + 0 * q4                       # lanes_of(.., fbG)
    b_lane = g_lane - q4                                            # g - fbB
    na_lane = (0x10003 - q4) - g_lane                               # fbNa - g
    for lane in (g_lane, b_lane, na_lane):                          # no borrow / carry between the two lanes of a word
        assert lane.min() >= 1 and lane.max() <= 0xFFFE
    G, B, NA = g_lane >> 15, b_lane >> 15, na_lane >> 15
    assert np.array_equal(G.astype(bool), (e >= 0) + 0 * q4)
    assert np.array_equal(B.astype(bool), e >= q4)
    assert np.array_equal(NA.astype(bool), e < 4 - q4)
    assert not np.any(NA & B) and not np.any(B & (1 - G))           # NA, B exclusive; B implies G
    # the reference's flags in terms of the three
    gez, lqt = e >= 0, (e >= 4 - q4) & (e < q4)
    assert np.array_equal((1 - G).astype(bool), ~gez + (0 * q4).astype(bool)) and np.array_equal(1 - NA - B, lqt.astype(np.int64))
    # DXT1: index = 2*Lqt + !Gez = 3 - (2 NA + 2 B + G): a row byte is 255 minus the weighted flags (weights 2*4^x and 4^x)
    assert np.array_equal(2 * lqt + (1 - gez.astype(np.int64)), 3 - (2 * NA + 2 * B + G))
    assert sum(3 * 4 ** x for x in range(4)) == 255
    assert [2 * 4 ** x for x in range(4)] == [0x02, 0x08, 0x20, 0x80] and [4 ** x for x in range(4)] == [0x01, 0x04, 0x10, 0x40]
    # ETC1s planes: pixel (x, y) sits at plane bit ((x^2)<<2)+y, i.e. byte (x<2), bit 4*(x&1)+y: weights 2^y and 2^(4+y)
    for x in range(4):
        for y in range(4):
            bit = ((x ^ 2) << 2) + y
            assert bit // 8 == (1 if x < 2 else 0) and bit % 8 == 4 * (x & 1) + y


def test_third_by_multiply():
    """floor(x / 3) == (x * 43691) >> 17 for every sum of three bytes (block_decode.cuh, BC1 thirds)."""
    x = np.arange(0, 766, dtype=np.int64)
    assert np.array_equal(x // 3, (x * 43691) >> 17)
