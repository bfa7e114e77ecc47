"""CPU model of the RaBitQ filter epilogue (qtc2_kernel<Q_RABITQ>, vecgo_b200/csrc/vg_quant_tc.cu) against the
certificate terms of qtc_exact_kernel.  The +-1 GEMM is exact (acc = D - 2 Hamming); the filter value is
    s' = fl(fl(f_q * acc + yn) * yn),  f_q = fl(fl(-2 qn) / D),  c_q = fl(qn qn)
and the reference estimator (rabitq.go:170-175, unfused float32) is (qn - yn)^2 + ((4 qn) yn / D) h.  The certificate
assumes |s' + c_q - reference| <= Eb + eref with Eb = smax (2^-21 + G 2^-23), eref = (qn + yn)^2 2^-21."""
import numpy as np

F = np.float32


def f(x):
    return np.asarray(x, np.float64).astype(F)


def test_rabitq_epilogue_stays_inside_the_certificate_terms():
    rng = np.random.default_rng(44)
    G = 128
    for D in (128, 768, 1536):
        for scale in (1e-3, 1.0, 300.0):
            n = 20000
            yn = f(np.abs(rng.standard_normal(n)) * scale * np.sqrt(D) * (0.5 + rng.random(n)))
            h = rng.binomial(D, 0.5, n).astype(np.int64)
            h[:50] = rng.integers(0, D + 1, 50)          # the whole range, including 0 and D
            for _ in range(3):
                qn = f(abs(rng.standard_normal()) * scale * np.sqrt(D) + 1e-6 * scale)
                acc = (D - 2 * h).astype(np.float64)     # exact in the tensor core's float32 accumulator
                fq = f(f(F(-2.0) * qn).astype(np.float64) / D)
                t = f(fq.astype(np.float64) * acc + yn.astype(np.float64))          # one FFMA
                s = f(t.astype(np.float64) * yn.astype(np.float64))                 # one FMUL
                cq = f(qn.astype(np.float64) * qn.astype(np.float64))
                # reference: t1 = qn - yn; a = ((4 qn) yn) / D; dist = t1 t1 + a h, every operation rounded to float32
                t1 = f(qn.astype(np.float64) - yn.astype(np.float64))
                a = f(f(f(F(4.0) * qn).astype(np.float64) * yn.astype(np.float64)).astype(np.float64) / D)
                ref = f(f(t1.astype(np.float64) ** 2).astype(np.float64) + f(a.astype(np.float64) * h).astype(np.float64))
                xx = float(yn.astype(np.float64).max()) ** 2
                qn_ = float(qn)
                y = np.sqrt(xx)
                smax = xx + 2.0 * qn_ * y
                Eb = smax * (1.0 / 2097152.0 + G / 8388608.0)
                eref = (qn_ + y) ** 2 / 2097152.0
                err = np.abs(s.astype(np.float64) + float(cq) - ref.astype(np.float64)).max()
                assert err <= Eb + eref, (D, scale, err, Eb + eref)
