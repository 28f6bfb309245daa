import numpy as np, sys
sys.path.insert(0, '/root/repo')
from smore_b200 import capi
from smore_b200 import dist as sdist
from tests.test_gpu_sharded import _sbm, _params
off, col, ww, test_s, test_d, train_adj = _sbm()
V, dim, total = len(off) - 1, 32, 12_000_000
init = ((np.random.default_rng(1).random((V, dim)) - 0.5) / dim)
world, rounds = 4, 20
gs, ms = [], []
for r in range(world):
    gr = capi.Graph.from_csr(off, col, ww); print(gr.set_shard(r, world))
    mr = capi.Model(gr, dim, 2, capi.F32)
    rows = sdist.owned_rows(V, r, world)
    mr.set_rows(0, init[rows]); mr.set_rows(1, np.zeros((len(rows), dim)))
    gs.append(gr); ms.append(mr)
for t in range(2):
    ptrs = [mr.device_ptr(t) for mr in ms]
    for mr in ms: mr.set_peer_ptrs(t, ptrs)
for k in range(3):
    for r, mr in enumerate(ms):
        p = _params(total // rounds, 100 + k); p.stream_base = r * (1 << 20)
        st = mr.train_line(p)
        a = mr.get_rows(0); c = mr.get_rows(1)
        print(k, r, st["samples"], "nan", np.isnan(a).sum(), np.isnan(c).sum(), "max", np.nanmax(np.abs(a)), np.nanmax(np.abs(c)), flush=True)
