"""The Q4 specification (oracle/kv_oracle.py q4_pack_chunk) re-derived in EXACT rational arithmetic, element by
element, with only the two roundings the spec names (scale -> bf16, 1/scale -> float32): the numpy oracle's float64
shortcut for "rint of the exact product" must agree everywhere, including ties and the extremes of a group."""
from fractions import Fraction

import numpy as np

from oracle import kv_oracle as ko


def rn_to(frac: Fraction, mant_bits: int) -> Fraction:
    """Round a positive rational to `mant_bits` significant bits, ties to even (normal range only)."""
    if frac == 0:
        return Fraction(0)
    e = 0
    while frac >= 2:
        frac /= 2
        e += 1
    while frac < 1:
        frac *= 2
        e -= 1
    scaled = frac * (1 << (mant_bits - 1))
    fl = scaled.numerator // scaled.denominator
    rem = scaled - fl
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and fl % 2 == 1):
        fl += 1
    return Fraction(fl, 1 << (mant_bits - 1)) * (Fraction(2) ** e)


def rint_even(frac: Fraction) -> int:
    fl = frac.numerator // frac.denominator
    rem = frac - fl
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and fl % 2 == 1):
        fl += 1
    return fl


def test_q4_codes_and_scales_equal_exact_rational_arithmetic():
    rng = np.random.default_rng(99)
    x = rng.standard_normal((1, 2, 6, 2, 64)).astype(np.float32) * np.float32(3.0)
    x[0, 0, 0, 0, :32] = 0.0                                   # an all-zero group
    x[0, 1, 1, 1, 32:] = np.float32(7.0) * np.float32(0.5)     # every element = absmax: codes +-7
    x[0, 1, 2, 0, :32] = np.arange(-16, 16, dtype=np.float32) * np.float32(0.4375)   # many exact .5 products
    bits = ko.f32_to_bf16_bits_rn(x)
    codes, sbits = ko.q4_pack_chunk(bits)
    xb = ko.bf16_bits_to_f32(bits)
    L, two, n, H, D = xb.shape
    G = ko.Q4_GROUP
    for idx in np.ndindex(L, two, n, H, D // G):
        grp = [Fraction(float(v)) for v in xb[idx[0], idx[1], idx[2], idx[3], idx[4] * G:(idx[4] + 1) * G]]
        amax = max(abs(v) for v in grp)
        s = rn_to(amax / 7, 8) if amax != 0 else Fraction(1)              # bf16: 8 significant bits
        assert Fraction(float(ko.bf16_bits_to_f32(np.array([sbits[idx]], np.uint16))[0])) == s, idx
        inv = rn_to(1 / s, 24)                                            # float32: 24 significant bits
        for j, v in enumerate(grp):
            q = max(-7, min(7, rint_even(v * inv)))
            d = idx[4] * G + j
            byte = int(codes[idx[0], idx[1], idx[2], idx[3], d // 2])
            nib = (byte >> 4) if d % 2 else (byte & 0xF)
            assert nib == (q & 0xF), (idx, j, float(v), q, nib)
