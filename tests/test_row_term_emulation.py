"""CPU check of the C2 row term's arithmetic (families.cuh::logit_fast / pool.cuh::logit_fast2_rep).

The constants and both lookup tables are read LITERALLY from the product source, compiled into a plain-C emulation of the
device code (same fma sequence, same bit manipulations) and compared with an 80-bit evaluation of
min(z, 0) - log1p(exp(-|z|)): a typo in a coefficient or a table entry fails here, without a GPU.  The device result
itself is checked on the GPU by test_gpu_parity.py::test_logit_term_accuracy_and_support."""
import ctypes
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "blangsdk_b200", "csrc", "families.cuh")

EMUL = r"""
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "tables.h"
static int lo32(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)u; }
static int hi32(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)(u >> 32); }
static double mk(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &u, 8); return d; }
double row_term(double z) {
  const double a = fabs(z);
  double kd = fma(a, kLgtK[0], kLgtK[1]);
  const int m = lo32(kd);
  kd -= kLgtK[1];
  const double te = kLgtExpTabC[m & 63];
  const double f = fma(a, kLgtK[0], -kd);
  double p = kLgtExp2C[0];
  for (int i = 1; i < 6; i++) p = fma(p, f, kLgtExp2C[i]);
  const double tp = te * p;
  const double t = mk(hi32(tp) + ((m >> 6) << 20), lo32(tp));
  const double u = 1.0 + t;
  unsigned j = (unsigned)(hi32(u) - 0x3ff00000) >> 12;
  if (j > 255u) j = 255u;
  const double rr = fma(u, kLgtLogTabC[j][0], -1.0);
  double q = kLgtLogC[0];
  for (int i = 1; i < 4; i++) q = fma(q, rr, kLgtLogC[i]);
  const double l = fma(rr, fma(rr, q, 1.0), kLgtLogTabC[j][1]);
  return fma(0.5, z - fabs(z), -l);
}
/* families.cuh::exp_tab / log_tab (the mixture kernels): Cody-Waite reduction, polynomial in r; table log */
double exp_tab(double x) {
  double kd = fma(x, -kLgtK[0], kLgtK[1]);
  const int m = lo32(kd);
  kd -= kLgtK[1];
  const double te = kLgtExpTabC[m & 63];
  double r = fma(kd, kLgtK[2], x);
  r = fma(kd, kLgtK[3], r);
  double p = kLgtExpC[0];
  for (int i = 1; i < 6; i++) p = fma(p, r, kLgtExpC[i]);
  const double tp = te * p;
  return mk(hi32(tp) + ((m >> 6) << 20), lo32(tp));
}
double log_tab(double a) {
  const int hi = hi32(a);
  const int e = (hi >> 20) - 1023;
  const double m = mk((hi & 0x000fffff) | 0x3ff00000, lo32(a));
  const unsigned j = ((unsigned)hi & 0x000ff000u) >> 12;
  const double rr = fma(m, kLgtLogTabC[j][0], -1.0);
  double q = kLgtLogC[0];
  for (int i = 1; i < 4; i++) q = fma(q, rr, kLgtLogC[i]);
  const double lm = fma(rr * rr, q, rr) + kLgtLogTabC[j][1];
  return fma((double)e, 6.93147180559945309417e-01, lm);
}
void exp_log_errors(const double* x, int n, double* rel_exp, double* abs_log) {   /* x <= 0: exp(x), log(exp-ish positive) */
  for (int i = 0; i < n; i++) {
    const long double ex = expl((long double)x[i]);
    rel_exp[i] = (double)(((long double)exp_tab(x[i]) - ex) / ex);
    const double a = (double)expl((long double)x[i] * 0.5L) * 3.7;      /* positive, exponents from -500 to 2 */
    abs_log[i] = (double)((long double)log_tab(a) - logl((long double)a));
  }
}
void row_terms(const double* z, int n, double* out, double* err) {
  for (int i = 0; i < n; i++) {
    out[i] = row_term(z[i]);
    const long double zz = z[i];
    const long double exact = fminl(zz, 0.0L) - log1pl(expl(-fabsl(zz)));
    err[i] = (double)((long double)out[i] - exact);
  }
}
"""


def _array(text, name):
    m = re.search(r"%s\[\d+\]\s*=\s*\{(.*?)\};" % name, text, re.S)
    assert m, name
    body = re.sub(r"//[^\n]*", "", m.group(1))
    return [float(x) for x in re.findall(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?", body)]


def _build(tmp_path):
    text = open(SRC).read()
    k, e2, lc = _array(text, "kLgtK"), _array(text, "kLgtExp2C"), _array(text, "kLgtLogC")
    ex, lg = _array(text, "kLgtExpTabC"), _array(text, "kLgtLogTabC")
    assert (len(k), len(e2), len(lc), len(ex), len(lg)) == (4, 6, 4, 64, 512)
    with open(tmp_path / "tables.h", "w") as f:
        for name, v in (("kLgtK", k), ("kLgtExp2C", e2), ("kLgtLogC", lc), ("kLgtExpTabC", ex), ("kLgtExpC", _array(text, "kLgtExpC"))):
            f.write("static const double %s[%d] = {%s};\n" % (name, len(v), ", ".join(repr(x) for x in v)))
        f.write("static const double kLgtLogTabC[256][2] = {%s};\n" % ", ".join("{%r, %r}" % (lg[2 * i], lg[2 * i + 1]) for i in range(256)))
    (tmp_path / "emul.c").write_text(EMUL)
    so = tmp_path / "emul.so"
    # -ffp-contract=off: only the fma() calls written above may fuse, as on the device
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(tmp_path / "emul.c"), "-lm"])
    lib = ctypes.CDLL(str(so))
    lib.row_terms.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.exp_log_errors.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return lib, (k, e2, ex, lg)


def test_tables_are_what_the_generator_defines(tmp_path):
    _, (k, e2, ex, lg) = _build(tmp_path)
    ln2 = np.log(np.longdouble(2))
    assert k[0] == float(-64 / ln2) and k[1] == 1.5 * 2.0 ** 52
    for i, c in enumerate(reversed(e2)):          # (ln2/64)^i / i!
        want = (ln2 / 64) ** i / np.longdouble(np.prod(np.arange(1, i + 1, dtype=np.float64)) if i else 1.0)
        assert abs(np.longdouble(c) - want) <= 0.51 * np.spacing(c), (i, c)      # every entry is the correctly rounded value
    for j in range(64):
        assert abs(np.longdouble(ex[j]) - np.longdouble(2) ** (np.longdouble(j) / 64)) <= 0.51 * np.spacing(ex[j]), j
    for j in range(256):
        inv, neg_log = lg[2 * j], lg[2 * j + 1]
        c = 1 + (np.longdouble(j) + np.longdouble(0.5)) / 256
        assert abs(np.longdouble(inv) - 1 / c) <= 0.51 * np.spacing(inv), j
        # the log of the value really multiplied by, not of 1/c_j
        assert abs(np.longdouble(neg_log) + np.log(np.longdouble(inv))) <= 0.51 * np.spacing(abs(neg_log)) + 1e-19, j


def test_row_term_emulation_against_extended_precision(tmp_path):
    lib, _ = _build(tmp_path)
    rng = np.random.default_rng(11)
    z = np.concatenate([rng.uniform(-12, 45, 400_000), rng.normal(0, 3, 400_000).clip(-12, None), rng.uniform(-12, 700, 100_000),
                        np.linspace(-12, 12, 48_001), [0.0, -0.0, 1e-300, -1e-300, 699.999, -12.0, 36.7, 36.8]])
    z = np.ascontiguousarray(z, dtype=np.float64)
    out, err = np.empty_like(z), np.empty_like(z)
    lib.row_terms(z.ctypes.data, z.size, out.ctypes.data, err.ctypes.data)
    assert np.all(np.isfinite(out)) and np.all(out <= 0.0)
    # absolute accuracy (the terms are summed): the table's and the polynomials' 1e-16 plus the rounding of the result
    assert np.all(np.abs(err) <= 2.5e-16 + np.spacing(np.abs(out)))
    assert np.max(np.abs(err)) < 1.1e-15                       # worst case is the output's own rounding near z = -12
    assert abs(np.mean(err)) < 5e-17                           # no systematic offset (sums of 1e5 rows stay below 1e-11)


def test_mixture_exp_and_log_emulation(tmp_path):
    """families.cuh::exp_tab / log_tab, the table-driven exp and log of the mixture kernels (C3), same emulation"""
    lib, (k, _, _, _) = _build(tmp_path)
    ln2 = np.log(np.longdouble(2))
    hi, lo = k[2], k[3]                       # -ln2/64 split: hi has 21 low mantissa bits clear, hi + lo = -ln2/64
    assert np.float64(hi).view(np.uint64) & np.uint64((1 << 21) - 1) == 0
    assert abs((np.longdouble(hi) + np.longdouble(lo)) + ln2 / 64) < 1e-19    # what an 80-bit sum can resolve
    rng = np.random.default_rng(12)
    x = np.ascontiguousarray(np.concatenate([rng.uniform(-700, 0, 300_000), rng.uniform(-40, 0, 300_000), [0.0, -700.0, -1e-300]]))
    rel, lg = np.empty_like(x), np.empty_like(x)
    lib.exp_log_errors(x.ctypes.data, x.size, rel.ctypes.data, lg.ctypes.data)
    assert np.max(np.abs(rel)) < 4e-16    # table entry, polynomial and product each round once: under two ulps of exp(x)
    assert np.max(np.abs(lg) / np.maximum(1.0, np.abs(0.5 * x + np.log(3.7)))) < 2.3e-16
