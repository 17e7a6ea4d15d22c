"""Small workload for ncu: one voxelisation + one align of the config-1 pair (pclomp / DIRECT7), then pclpca / DIRECT1."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import lv_slam_b200 as L
from lv_slam_b200 import synth
tgt, src, guess, truth = synth.config1_pair()
for variant, search, iters in ((0, 2, 3), (1, 3, 3)):
    n = L.NormalDistributionsTransform(variant=variant)
    n.setTransformationEpsilon(0.01); n.setMaximumIterations(iters); n.setNeighborhoodSearchMethod(search)
    n.setInputTarget(tgt); n.setInputSource(src)
    n.align(guess)
    print(n.result()["iterations"])
