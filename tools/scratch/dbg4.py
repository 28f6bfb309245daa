import numpy as np, sys
sys.path.insert(0, '/root/repo')
from smore_b200 import capi
from tests.test_gpu_sharded import _sbm, _params
from tests.test_gpu_quality import evaluate
off, col, ww, test_s, test_d, train_adj = _sbm()
V, dim = len(off) - 1, 32
for total in (3_000_000, 6_000_000, 12_000_000):
    g = capi.Graph.from_csr(off, col, ww)
    m = capi.Model(g, dim, 2, capi.F32)
    m.init(0, True, 5); m.init(1, False, 5)
    m.train_line(_params(total, 17))
    a, r = evaluate(m.get_rows(0), m.get_rows(1), test_s, test_d, train_adj, np.random.default_rng(2))
    print(total, "AUC", a, "rec", r, flush=True)
