import sys, time
sys.path.insert(0, ".")
from kamphir_b200 import PhyloKernel, synth
tips = int(sys.argv[1]); nreps = int(sys.argv[2])
pk = PhyloKernel(decayFactor=0.2, gaussFactor=2.0, normalize="mean")
obs = synth.kingman_newick(tips, 7)
sims = [synth.kingman_newick(tips, 2000 + i).encode() for i in range(nreps)]
from kamphir_b200 import FlatTree
fo = FlatTree.from_newick(obs)
for _ in range(6):
    t0 = time.perf_counter()
    r = pk.score_batch(fo, sims, use_kcoal=True)
    print("python call %.0f us" % (1e6 * (time.perf_counter() - t0)), file=sys.stderr)
