"""Regenerates the numeric constants embedded in real-time-sdr_b200/csrc/pllmath.cuh (needs mpmath).

Run:  python tests/gen_pllmath_consts.py      # prints the constants; compare with the header
"""
import mpmath as mp

mp.mp.prec = 400


def dbl(x):
    return float(mp.mpf(x))


def split_bits(x, bits):
    x = mp.mpf(x)
    e = mp.floor(mp.log(abs(x), 2))
    scale = mp.mpf(2) ** (e - bits + 1)
    return mp.floor(x / scale) * scale


def triple(x):
    a = dbl(x)
    b = dbl(mp.mpf(x) - mp.mpf(a))
    c = dbl(mp.mpf(x) - mp.mpf(a) - mp.mpf(b))
    return a, b, c


def remez(g, n, lo, hi, w, iters=12, grid_n=4000):
    """Weighted minimax polynomial of degree n-1 in z for g on [lo, hi] (plain exchange algorithm on a fine grid)."""
    xs = [(lo + hi) / 2 + (hi - lo) / 2 * mp.cos(mp.pi * k / n) for k in range(n + 1)][::-1]
    c = None
    for _ in range(iters):
        A, b = mp.matrix(n + 1, n + 1), mp.matrix(n + 1, 1)
        for i, x in enumerate(xs):
            for j in range(n):
                A[i, j] = x ** j
            A[i, n] = (-1) ** i / w(x)
            b[i] = g(x)
        sol = mp.lu_solve(A, b)
        c = [sol[j] for j in range(n)]
        grid = [lo + (hi - lo) * k / grid_n for k in range(grid_n + 1)]
        ev = [(sum(c[j] * x ** j for j in range(n)) - g(x)) * w(x) for x in grid]
        ext = [(grid[k], ev[k]) for k in range(grid_n + 1)
               if (k == 0 or abs(ev[k]) >= abs(ev[k - 1])) and (k == grid_n or abs(ev[k]) >= abs(ev[k + 1]))]
        pts = []
        for x, v in ext:
            if pts and mp.sign(pts[-1][1]) == mp.sign(v):
                if abs(v) > abs(pts[-1][1]):
                    pts[-1] = (x, v)
            else:
                pts.append((x, v))
        if len(pts) < n + 1:
            break
        while len(pts) > n + 1:
            pts.pop(0) if abs(pts[0][1]) < abs(pts[-1][1]) else pts.pop()
        xs = [q[0] for q in pts]
    return c


def lean_kernels():
    """SDRB_PLL_LEAN_CONSTS: sin r = r + r z (L1 + .. + L5 z^4), cos r = 1 - z/2 + z^2 (M1 + .. + M5 z^4), z = r^2, |r| <= pi/4."""
    Z = (mp.pi / 4) ** 2
    tiny = mp.mpf("1e-40")

    def gs(z):
        r = mp.sqrt(z)
        return (mp.sin(r) / r - 1) / z if z else mp.mpf(-1) / 6

    def gc(z):
        r = mp.sqrt(z)
        return (mp.cos(r) - 1 + z / 2) / (z * z) if z else mp.mpf(1) / 24

    L = remez(gs, 5, mp.mpf(0), Z, lambda z: z if z else tiny)       # error relative to sin r ~ r
    M = remez(gc, 5, mp.mpf(0), Z, lambda z: z * z if z else tiny)   # absolute error of cos r
    print("lean sin L1..L5:", [dbl(x).hex() for x in L])
    print("lean cos M1..M5:", [dbl(x).hex() for x in M])


def main():
    lean_kernels()
    pio2 = mp.pi / 2
    rem, parts = pio2, []
    for _ in range(5):
        p = split_bits(rem, 22)
        parts.append(p)
        rem -= p
    for i, p in enumerate(parts):
        print(f"kP{i + 1} = {float(p).hex()}")
    t1 = dbl(rem)
    print("kPTailHi =", t1.hex(), " kPTailLo =", dbl(rem - mp.mpf(t1)).hex())
    print("kP4Rest =", dbl(pio2 - parts[0] - parts[1] - parts[2]).hex())
    print("kTwoOverPi =", dbl(2 / mp.pi).hex())
    for name, v in (("kPi", mp.pi), ("kPio2", pio2)):
        print(name, [x.hex() for x in triple(v)])
    print("sin Taylor S1..S8:", [dbl(mp.mpf((-1) ** j) / mp.factorial(2 * j + 1)).hex() for j in range(1, 9)])
    print("cos Taylor C2..C9:", [dbl(mp.mpf((-1) ** j) / mp.factorial(2 * j)).hex() for j in range(2, 10)])
    print("atan Taylor A1..A5:", [dbl(mp.mpf((-1) ** j) / (2 * j + 1)).hex() for j in range(1, 6)])
    print("atan(i/16) hi:", [triple(mp.atan(mp.mpf(i) / 16))[0].hex() for i in range(17)])
    print("atan(i/16) lo:", [triple(mp.atan(mp.mpf(i) / 16))[1].hex() for i in range(17)])
    print("(float)pi, (float)(pi/2):", repr(float(mp.mpf(mp.pi))), repr(float(mp.pi / 2)))


if __name__ == "__main__":
    main()
