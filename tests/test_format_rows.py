"""CPU: the embedding writer's number formatting (smore_format_rows, host-only) against printf semantics -- the
reference writes `ostream << double` (== "%g", src/model/LINE.cpp:35-38) or Go "%.6f" (line.go:226)."""
import numpy as np

from smore_b200 import capi


def _expected(rows, first, fmt):
    spec = "%g" if fmt == 0 else "%.6f"
    return "".join(str(first + r) + "".join(" " + spec % float(x) for x in row) + "\n" for r, row in enumerate(rows)).encode()


def test_format_matches_printf():
    rng = np.random.default_rng(7)
    n, dim = 3000, 17
    rows = (rng.random((n, dim)) - 0.5) / 64
    rows[0, :8] = [0.0, -0.0, 1.0, -1.5, 1e-5, 9.99995e-5, 123456.5, 1234567.0]
    rows[1, :8] = [1e-7, -3.25e-12, 0.1, 0.0000005, 0.0000015, 2.5e-6, 1e100, -1e-300]
    rows[2] = rng.standard_normal(dim) * 1e3
    rows[3] = np.float32(rng.standard_normal(dim)).astype(np.float64)  # values as an fp32 table holds them
    for fmt in (0, 1):
        got = capi.format_rows(rows, first_id=41, fmt=fmt)
        assert got == _expected(rows, 41, fmt)


def test_format_empty_and_threads_agree():
    assert capi.format_rows(np.zeros((0, 4))) == b""
    rows = np.arange(12, dtype=np.float64).reshape(3, 4) / 7
    assert capi.format_rows(rows, fmt=1) == _expected(rows, 0, 1)


def test_format_fast_paths_on_many_magnitudes():
    """The fast paths (one-rounding decimal scaling, fall-back near rounding boundaries) against printf on values of
    every magnitude they accept and just outside, on dyadic fractions (exact ties), on 6-digit decimals +- a few ulps
    (boundary cases) and on fp32-representable values."""
    rng = np.random.default_rng(11)
    vals = [np.ldexp(rng.random(200_000) + 0.5, rng.integers(-40, 40, 200_000)) * rng.choice([-1.0, 1.0], 200_000)]
    vals.append(10.0 ** rng.uniform(-9, 9, 200_000) * rng.choice([-1.0, 1.0], 200_000))
    vals.append(rng.integers(-4096, 4096, 50_000) / 2.0 ** rng.integers(0, 24, 50_000))          # dyadic: exact ties
    dec = rng.integers(100_000, 1_000_000, 100_000) * 10.0 ** rng.integers(-12, 4, 100_000)       # d.ddddd x 10^k
    half = (rng.integers(100_000, 1_000_000, 100_000) + 0.5) * 10.0 ** rng.integers(-12, 4, 100_000)
    for base in (dec, half):
        for ulps in (-2, -1, 0, 1, 2):
            v = base.copy()
            for _ in range(abs(ulps)):
                v = np.nextafter(v, np.inf if ulps > 0 else -np.inf)
            vals.append(v)
    vals.append(np.float32((rng.random(200_000) - 0.5) / 128).astype(np.float64))
    vals.append(np.array([0.0, -0.0, 1e-7, 9.999995e-5, 0.0001, 999999.5, 999999.4999999999, 1e6, 1e7, 9999999.5, 4095.9999995,
                          4096.0, 0.5, 0.0078125, 1.5e-6, 2.5e-6, 123456.5, 1e-5, 99999.95, 0.9999995, 0.99999949999]))
    x = np.concatenate(vals)
    x = x[: len(x) // 16 * 16].reshape(-1, 16)
    for fmt in (0, 1):
        got = capi.format_rows(x, first_id=0, fmt=fmt).split(b"\n")
        spec = "%g" if fmt == 0 else "%.6f"
        for r in range(x.shape[0]):
            exp = (str(r) + "".join(" " + spec % float(v) for v in x[r])).encode()
            assert got[r] == exp, (fmt, r, x[r].tolist(), got[r], exp)
