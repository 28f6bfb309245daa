import numpy as np, sys
sys.path.insert(0, '/root/repo')
from oracle import bindings as B
from smore_b200 import capi
from smore_b200 import dist as sdist
from tests import graphs
src, dst, w = graphs.random_graph(300, 4000, seed=61)
off, col, ww, _ = B.edges_to_csr(src, dst, w, 1)
V=len(off)-1; dim=16; world=2
init = ((np.random.default_rng(1).random((V, dim)) - 0.5) / dim)
gs, ms = [], []
for r in range(world):
    g = capi.Graph.from_csr(off, col, ww); print(g.set_shard(r, world))
    m = capi.Model(g, dim, 2, capi.F64)
    rows = sdist.owned_rows(V, r, world)
    m.set_rows(0, init[rows]); m.set_rows(1, np.zeros((len(rows), dim)))
    gs.append(g); ms.append(m)
for t in range(2):
    ptrs = [m.device_ptr(t) for m in ms]
    print("ptrs", t, [hex(p) for p in ptrs])
    for m in ms: m.set_peer_ptrs(t, ptrs)
p = capi.default_params(); p.mode = capi.MODE_DETERMINISTIC; p.seed=3; p.total=4000
st = ms[0].train_line(p); print(st)
for r in range(world):
    a = ms[r].get_rows(0); c = ms[r].get_rows(1)
    rows = sdist.owned_rows(V, r, world)
    print("rank", r, "Wv rows changed", (np.abs(a-init[rows]).max(1)>0).sum(), "Wc rows changed", (np.abs(c).max(1)>0).sum(), "max", np.abs(a).max(), np.abs(c).max())
# emulate rank 0's run in numpy? compare pos scores
