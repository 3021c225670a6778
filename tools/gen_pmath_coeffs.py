#!/usr/bin/env python
"""Generate the polynomial coefficients / split constants used by sosie_b200/csrc/pmath.h.

Everything is derived from first principles with mpmath (60 digits): Chebyshev-node
interpolation of the residual functions, converted to monomial coefficients.  The output is
pasted into pmath.h as C hex-float literals, so the header itself has no generator
dependency.  Run:  python tools/gen_pmath_coeffs.py > /tmp/pmath_coeffs.txt
"""
import mpmath as mp

mp.mp.dps = 80


def cheb_fit(f, a, b, deg):
    """Monomial coefficients (in t, t in [a,b]) of the degree-`deg` Chebyshev interpolant of f."""
    n = deg + 1
    # nodes in [-1,1]
    xs = [mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    ts = [(a + b) / 2 + (b - a) / 2 * x for x in xs]
    fs = [f(t) for t in ts]
    # Chebyshev coefficients
    c = []
    for j in range(n):
        s = mp.mpf(0)
        for k in range(n):
            s += fs[k] * mp.cos(mp.pi * j * (2 * k + 1) / (2 * n))
        c.append(2 * s / n)
    c[0] /= 2
    # convert sum c_j T_j(x) to monomial in x
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, n):
        prev, prev2 = T[j - 1], T[j - 2]
        cur = [mp.mpf(0)] + [2 * v for v in prev]
        for i, v in enumerate(prev2):
            cur[i] -= v
        T.append(cur)
    px = [mp.mpf(0)] * n
    for j in range(n):
        for i, v in enumerate(T[j]):
            px[i] += c[j] * v
    # x = (t - m)/h  ->  monomial in t
    m = (a + b) / 2
    h = (b - a) / 2
    # expand sum px_i ((t-m)/h)^i
    pt = [mp.mpf(0)] * n
    for i in range(n):
        # ((t-m)/h)^i = sum_k C(i,k) t^k (-m)^(i-k) / h^i
        for k in range(i + 1):
            pt[k] += px[i] * mp.binomial(i, k) * (-m) ** (i - k) / h ** i
    return pt


def maxerr(f, coeffs, a, b, n=4001, rel=False):
    e = mp.mpf(0)
    for k in range(n):
        t = a + (b - a) * mp.mpf(k) / (n - 1)
        p = mp.mpf(0)
        for cc in reversed(coeffs):
            p = p * t + cc
        d = abs(p - f(t))
        if rel:
            ft = abs(f(t))
            if ft > 0:
                d /= ft
        e = max(e, d)
    return e


def hexd(x):
    return float(x).hex()


def emit(name, coeffs):
    print(f"// {name}")
    for i, cc in enumerate(coeffs):
        print(f"  {name}[{i}] = {hexd(cc)}  // {mp.nstr(cc, 20)}")


def split_const(x, bits_list):
    """Split x into pieces with the given mantissa widths (truncation), last piece = remainder rounded."""
    out = []
    rem = mp.mpf(x)
    for bits in bits_list:
        e = mp.floor(mp.log(abs(rem), 2))
        q = mp.mpf(2) ** (e - bits + 1)
        piece = mp.floor(rem / q) * q
        out.append(piece)
        rem -= piece
    out.append(rem)
    return out


if __name__ == "__main__":
    # ---- sin kernel on |x| <= pi/4 (a little wider):  sin(x) = x + x^3 * S(z), z = x^2
    zmax = (mp.pi / 4 * mp.mpf("1.02")) ** 2

    def fS(z):
        if z == 0:
            return -mp.mpf(1) / 6
        x = mp.sqrt(z)
        return (mp.sin(x) - x) / (x * z)

    S = cheb_fit(fS, mp.mpf(0), zmax, 7)
    print("// sin S(z): max abs err", mp.nstr(maxerr(fS, S, 0, zmax), 5))
    emit("S", S)

    # ---- cos kernel: cos(x) = 1 - z/2 + z^2 * C(z)
    def fC(z):
        if z == 0:
            return mp.mpf(1) / 24
        x = mp.sqrt(z)
        return (mp.cos(x) - 1 + z / 2) / (z * z)

    C = cheb_fit(fC, mp.mpf(0), zmax, 7)
    print("// cos C(z): max abs err", mp.nstr(maxerr(fC, C, 0, zmax), 5))
    emit("C", C)

    # ---- pi/2 split: 3 pieces of 33 bits + tail
    p = split_const(mp.pi / 2, [33, 33, 33])
    print("// pio2 split (33,33,33 bits + tail)")
    for i, v in enumerate(p):
        print(f"  PIO2_{i+1} = {hexd(v)}  // {mp.nstr(v, 25)}")
    print(f"  INVPIO2 = {hexd(2/mp.pi)}")
    # pi hi/lo
    pi_hi = mp.mpf(float(mp.pi))
    print(f"  PI_HI = {hexd(pi_hi)}  PI_LO = {hexd(mp.pi - pi_hi)}")
    pio2_hi = mp.mpf(float(mp.pi / 2))
    print(f"  PIO2_HI = {hexd(pio2_hi)}  PIO2_LO = {hexd(mp.pi/2 - pio2_hi)}")
    pio4_hi = mp.mpf(float(mp.pi / 4))
    print(f"  PIO4_HI = {hexd(pio4_hi)}  PIO4_LO = {hexd(mp.pi/4 - pio4_hi)}")

    # ---- asin kernel: asin(x) = x + x*z*R(z), z = x^2 in [0, 0.25]
    def fR(z):
        if z == 0:
            return mp.mpf(1) / 6
        x = mp.sqrt(z)
        return (mp.asin(x) - x) / (x * z)

    R = cheb_fit(fR, mp.mpf(0), mp.mpf("0.2501"), 17)
    print("// asin R(z): max abs err", mp.nstr(maxerr(fR, R, 0, mp.mpf("0.2501")), 5))
    emit("R", R)

    # ---- atan kernel: atan(x) = x + x*z*A(z), z = x^2, |x| <= 7/16
    za = (mp.mpf(7) / 16) ** 2 * mp.mpf("1.001")

    def fA(z):
        if z == 0:
            return -mp.mpf(1) / 3
        x = mp.sqrt(z)
        return (mp.atan(x) - x) / (x * z)

    A = cheb_fit(fA, mp.mpf(0), za, 13)
    print("// atan A(z): max abs err", mp.nstr(maxerr(fA, A, 0, za), 5))
    emit("A", A)
    for nm, v in (("ATAN_0_5", mp.atan(mp.mpf("0.5"))), ("ATAN_1_0", mp.atan(1)),
                  ("ATAN_1_5", mp.atan(mp.mpf("1.5"))), ("ATAN_INF", mp.pi / 2)):
        hi = mp.mpf(float(v))
        print(f"  {nm}_HI = {hexd(hi)}  {nm}_LO = {hexd(v - hi)}")

    # ---- log kernel: log(1+f) = 2s + s*L(z)  with s=f/(2+f), z=s^2;  L(z)=(log((1+s)/(1-s))-2s)/s
    smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1) * mp.mpf("1.001")   # f in [sqrt2/2-1, sqrt2-1]
    zl = smax ** 2

    def fL(z):
        # returns L(z)/z  so that log(1+f) = 2s + s*z*Lr(z)
        if z == 0:
            return mp.mpf(2) / 3
        s = mp.sqrt(z)
        return (mp.log((1 + s) / (1 - s)) - 2 * s) / (s * z)

    L = cheb_fit(fL, mp.mpf(0), zl, 8)
    print("// log Lr(z): max abs err", mp.nstr(maxerr(fL, L, 0, zl), 5))
    emit("L", L)
    ln2 = mp.log(2)
    l2 = split_const(ln2, [32])
    print(f"  LN2_HI = {hexd(l2[0])}  LN2_LO = {hexd(l2[1])}")

    # ---- expf kernel (double arithmetic): exp(r), |r| <= ln2/2
    rm = mp.log(2) / 2 * mp.mpf("1.01")
    E = cheb_fit(lambda r: mp.exp(r), -rm, rm, 9)
    print("// exp E(r): max rel err", mp.nstr(maxerr(lambda r: mp.exp(r), E, -rm, rm, rel=True), 5))
    emit("E", E)
    print(f"  LN2 = {hexd(ln2)}  INVLN2 = {hexd(1/ln2)}")
