"""Small mixed workload for compute-sanitizer (memcheck): ragged pairs on both C-storage paths, a Gram part, fp32."""
import sys
sys.path.insert(0, ".")
import numpy as np
from kamphir_b200 import FlatTree, PhyloKernel, get_context, synth

nwk = [synth.kingman_newick(n, 10 + n) for n in (2, 3, 5, 17, 33, 64, 100, 130)] + [synth.caterpillar_newick(40, 1),
                                                                                   synth.yule_newick(300, 5)]
flats = [FlatTree.from_newick(s) for s in nwk]
for prec in (64, 32):
    pk = PhyloKernel(decayFactor=0.2, gaussFactor=2.0, precision=prec)
    for storage in (0, 1, 2):
        get_context().configure(0, 0, storage)
        k = pk.kernel_batch(flats[:9])
        assert np.all(np.isfinite(k))
    get_context().configure(0, 0, 0)
    g = pk.gram(flats, normalised=True)
    assert np.allclose(np.diag(g), 1.0)
    r = pk.score_batch(flats[6], flats[:6])
print("sanitize case ok")
