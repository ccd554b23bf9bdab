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


def main():
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
