"""Derive exact integer formulas for vrsqrt14ps / vrcp14ps from their value tables.

    gcc -O2 -mavx512f -o gen tools/gen_svml14_tables.c && ./gen tables.bin     # needs an AVX-512F CPU
    python tools/fit_svml14_seeds.py tables.bin 360-to-planer-images_b200/csrc/svml14_tables.inc

Both instructions turn out to be piecewise linear with truncation in the top mantissa bits of their input
(64 segments for rcp14, 2 x 32 for rsqrt14): value = (base + ((rem - B * lo) >> 10)) << 7.  The script finds
the coefficients with a small LP per segment, rounds them to integers and verifies all 2 x 65536 values.
"""
import sys

import numpy as np
from scipy.optimize import linprog


def fit(v, hbits):
    seg = len(v) >> hbits
    out = []
    for h in range(1 << hbits):
        y = v[h * seg:(h + 1) * seg].astype(float)
        lo = np.arange(seg, dtype=float)
        A_ub = np.concatenate([np.stack([-np.ones(seg), lo, np.ones(seg)], 1),
                               np.stack([np.ones(seg), -lo, np.ones(seg)], 1)])
        b_ub = np.concatenate([-y, y + 1])
        r = linprog([0, 0, -1], A_ub=A_ub, b_ub=b_ub, bounds=[(None, None), (None, None), (0, None)], method="highs")
        if r.status != 0:
            raise SystemExit(f"segment {h} is not linear-with-truncation at {hbits} segment bits")
        out.append(r.x)
    return np.array(out)


def integerise(v, hbits, coef, s=10):
    seg = len(v) >> hbits
    A = np.zeros(1 << hbits, np.int64)
    B = np.zeros(1 << hbits, np.int64)
    lo = np.arange(seg)
    for h, (a, b, _m) in enumerate(coef):
        y = v[h * seg:(h + 1) * seg]
        B0, A0 = int(round(b * (1 << s))), int(np.floor(a * (1 << s)))
        for dB in range(-3, 4):
            hit = [dA for dA in range(-40, 41) if np.array_equal((A0 + dA - (B0 + dB) * lo) >> s, y)]
            if hit:
                A[h], B[h] = A0 + hit[0], B0 + dB
                break
        else:
            raise SystemExit(f"no integer coefficients for segment {h}")
    return A, B


def main(src, dst):
    tab = np.fromfile(src, dtype=np.uint32).astype(np.int64)
    assert tab.size == 131072 and not (tab & 127).any()
    rows = []
    for name, v, hb in (("rsq0", tab[:32768] >> 7, 5), ("rsq1", tab[32768:65536] >> 7, 5), ("rcp", tab[65536:] >> 7, 6)):
        A, B = integerise(v, hb, fit(v, hb))
        base, rem = A >> 10, A & 1023
        seg = len(v) >> hb
        lo = np.arange(seg)
        for h in range(1 << hb):
            assert np.array_equal(base[h] + ((rem[h] - B[h] * lo) >> 10), v[h * seg:(h + 1) * seg]), (name, h)
            rows += [int(base[h]), (int(rem[h]) << 16) | int(B[h])]
        print(name, "ok:", 1 << hb, "segments")
    with open(dst, "w") as f:
        f.write("// vrsqrt14ps / vrcp14ps (AVX-512F) as exact integer formulas, derived from the value tables that\n"
                "// tools/gen_svml14_tables.c dumps and verifies on an AVX-512F CPU (tools/fit_svml14_seeds.py).\n"
                "// Both instructions are piecewise linear with truncation in the top mantissa bits of the input:\n"
                "//   value_bits = (base + ((rem - B * lo) >> 10)) << 7,  lo = low 10 bits of the index,\n"
                "// one {base, rem << 16 | B} pair per segment: 32 + 32 segments for rsqrt14 (exponent parity 0 / 1,\n"
                "// index = top 15 mantissa bits), 64 for rcp14 (index = top 16 mantissa bits).  Exact powers are exact.\n")
        for i in range(0, len(rows), 8):
            f.write(",".join("0x%08xu" % x for x in rows[i:i + 8]) + ",\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
