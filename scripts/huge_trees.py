import sys, time
sys.path.insert(0, ".")
import numpy as np
from kamphir_b200 import FlatTree, PhyloKernel, get_context, synth
from oracle import c_oracle, phylok_oracle as po
sys.setrecursionlimit(100000)
pk = PhyloKernel(decayFactor=0.2, gaussFactor=2.0)
a, b = synth.kingman_newick(12000, 1), synth.yule_newick(9000, 2)
fa, fb = FlatTree.from_newick(a), FlatTree.from_newick(b)
t0 = time.time(); k = pk.kernel_batch([fa, fb], pairs=[(0, 1), (0, 0)]); t1 = time.time()
print("12000 x 9000 tips:", k, get_context().last_stats(), "%.2f s" % (t1 - t0), flush=True)
oa = c_oracle.CFlat(po.prepare(po.read_newick(a), "kamphir:mean")); ob = c_oracle.CFlat(po.prepare(po.read_newick(b), "kamphir:mean"))
w = c_oracle.kernel(oa, ob, 0.2, 2.0)
print("oracle", w, "rel err", abs(k[0] - w) / w, flush=True)
c = synth.kingman_newick(40000, 3)
fc = FlatTree.from_newick(c)
t0 = time.time(); kc = pk.kernel_batch([fc, fa], pairs=[(0, 0), (0, 1), (1, 0)]); t1 = time.time()
print("40000 tips self / vs 12000:", kc, get_context().last_stats(), "%.2f s" % (t1 - t0))
assert np.all(np.isfinite(kc)) and abs(kc[1] - kc[2]) <= 1e-9 * kc[1]
