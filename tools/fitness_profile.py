"""getFitnessScore on the config-1 pair, a few calls; run under ncu for the per-kernel times:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fit_launches.csv python tools/fitness_profile.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import lv_slam_b200 as L
from lv_slam_b200 import synth

tgt, src, guess, truth = synth.config1_pair()
n = L.NormalDistributionsTransform()
n.setInputTarget(tgt); n.setInputSource(src)
big = float(np.finfo(np.float64).max)
for T in (truth, truth, guess):
    print(n.getFitnessScore(big, T=T, with_count=True))
