# generates the exp / log tables used by logit_fast (50-digit arithmetic, rounded once to double)
from decimal import Decimal, getcontext
import struct
getcontext().prec = 60
def d2f(x): return float(x)   # Decimal -> nearest double (Python rounds correctly)
NE, NL = 64, 256
print("// 2^(j/64), j = 0..63")
rows = []
for j in range(NE):
    v = Decimal(2) ** (Decimal(j) / NE)
    rows.append(repr(d2f(v)))
print(", ".join(rows))
print("// {fl(1/c_j), -log(fl(1/c_j))}, c_j = 1 + (j + 1/2)/NL")
rows = []
for j in range(NL):
    c = Decimal(1) + (Decimal(j) + Decimal("0.5")) / NL
    inv = d2f(Decimal(1) / c)
    lc = -(Decimal(inv).ln())
    rows.append("{%r, %r}" % (inv, d2f(lc)))
print(", ".join(rows))
ln2 = Decimal(2).ln()
# Cody-Waite split of ln2/64: hi has its low 11+6 bits zero so m*hi is exact for |m| < 2^17
hi = d2f(ln2 / NE)
b = struct.unpack("<Q", struct.pack("<d", hi))[0] & ~((1 << 21) - 1)
hi = struct.unpack("<d", struct.pack("<Q", b))[0]
lo = d2f(ln2 / NE - Decimal(hi))
print("// -64*log2(e), ln2/64 hi, lo")
print(repr(d2f(-Decimal(NE) / ln2)), repr(hi), repr(lo))
# exp(-a) = 2^((m + f)/64) with f = a * (-64 log2 e) - m taken by ONE fma (no Cody-Waite pair): the polynomial is in f,
# coefficients (ln2/64)^i / i!, i = 5 .. 0
print("// (ln2/64)^i / i!, i = 5 .. 0")
fact = 1
cs = []
for i in range(6):
    cs.append(repr(d2f((ln2 / NE) ** i / fact)))
    fact *= (i + 1)
print(", ".join(reversed(cs)))
