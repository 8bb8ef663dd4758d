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
    for stage in (1, 2):      # tree parts staged in global memory: both, COLUMN part only
        get_context().set_option("unstaged", stage)
        k = pk.kernel_batch(flats[:9])
        assert np.all(np.isfinite(k))
        assert np.all(np.isfinite(pk.gram(flats[3:], normalised=False)))
    get_context().set_option("unstaged", 0)
    g = pk.gram(flats, normalised=True)
    assert np.allclose(np.diag(g), 1.0)
    r = pk.score_batch(flats[6], flats[:6])
    # labelled trees (chunk-dealing kernel variant), a CTA narrower than the rectangles (second chunk per warp), lambda = 0
    lab = [FlatTree.from_newick(synth.kingman_newick(90, 40 + i, states=3), 3, "mean", 3) for i in range(4)]
    kl = PhyloKernel(decayFactor=0.2, gaussFactor=2.0, precision=prec, n_states=3).kernel_batch(lab)
    assert np.all(np.isfinite(kl))
    get_context().configure(64, 0, 0)
    k2 = pk.kernel_batch(flats[5:])
    assert np.all(np.isfinite(k2))
    get_context().configure(0, 0, 0)
    assert np.all(PhyloKernel(decayFactor=0.0, precision=prec).kernel_batch(flats[:4]) == 0.0)
print("sanitize case ok")
