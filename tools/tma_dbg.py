import sys, numpy as np
sys.path.insert(0, '/root/repo')
from tests import decks
from epoch_b200.pic import Simulation
dk = decks.laser2d(n=64)
sim = Simulation(dk)
sim.init()
sim.fields_half()
sim.synchronize()
print("ok", float(np.abs(sim.download_field("bz")).max()))
