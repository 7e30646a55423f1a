import sys, json
import os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import numeric_gate as g

class TP(g.TwoPassTransform):
    def __init__(self, tf1, tf2, ti2, ti1):
        super().__init__(2)
        n1, n2 = np.arange(g.N1), np.arange(g.N2)
        W64 = np.exp(-2j * np.pi * np.outer(n1, n1) / g.N1)
        W32 = np.exp(-2j * np.pi * np.outer(n2, n2) / g.N2)
        self.f1, self.f2 = g.SplitDft(W64, tf1), g.SplitDft(W32, tf2)
        self.i1, self.i2 = g.SplitDft(W64.conj(), ti1), g.SplitDft(W32.conj(), ti2)

iters, T = int(sys.argv[1]), int(sys.argv[2])
cfgs = {"1111": (1,1,1,1), "2211": (2,2,1,1), "1122": (1,1,2,2), "2111": (2,1,1,1), "1211": (1,2,1,1)}
trs = {k: TP(*v) for k, v in cfgs.items()}
ref_tr = g.Fp64Transform()
for seed in (0, 1):
    rs = np.random.RandomState(seed)
    fams = {"uni+100": g.magnitudes(rs.rand(T, 1025), 100.0), "uni-100": g.magnitudes(rs.rand(T, 1025), -100.0),
            "speech": np.abs(g.stft(g.speechlike(g.HOP * (T - 1), 10 + seed), ref_tr))}
    for fam, S in fams.items():
        ang = np.exp(2j * np.pi * rs.rand(*S.shape))
        ref = g.griffin_lim(S, ang, iters, ref_tr)
        row = {k: round(g.snr_db(g.griffin_lim(S, ang, iters, tr), ref), 1) for k, tr in trs.items()}
        print(seed, fam, row, flush=True)
