import numpy as np, os, sys
sys.path.insert(0, '/root/repo')
from oracle import bindings as B
from smore_b200 import capi
G = np.load('/root/repo/tests/golden/golden_v1.npz')
SEED=20261018
off, col, ww, _ = B.edges_to_csr(G["g300_src"], G["g300_dst"], G["g300_w"], 1)
og = B.OracleGraph(B.SEM_CPP, off, col, ww)
dg = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP)
for order in (1,2):
  for total in (1000, 10000, 100000, 300000, 1000000):
    a, c = G["g300_init_v"].copy(), G["g300_init_c"].copy()
    pos = og.train_line_cpp(a, a if order==1 else c, 5, 0.025, total, SEED, 0)
    m = capi.Model(dg, 8, 1 if order==1 else 2, capi.F64)
    m.set_rows(0, G["g300_init_v"])
    if order==2: m.set_rows(1, G["g300_init_c"])
    p = capi.default_params(); p.semantics=0; p.mode=0; p.seed=SEED; p.total=total; p.order=order
    st = m.train_line(p)
    d = m.get_rows(0)
    print(order, total, st["words_stream0"]==pos, np.abs(d-a).max()/np.abs(a).max(), np.abs(a).max(), flush=True)
