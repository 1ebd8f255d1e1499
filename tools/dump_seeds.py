import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sapcu_b200, sapcu_b200.synthetic as syn
from sapcu_b200.generation import Generator3D6
gen = Generator3D6.__new__(Generator3D6)
gen.device, gen.dense_spacing = torch.device("cuda:0"), 0.004
cloud = syn.cloud(2048, seed=0, shape="sphere")
s = gen.gpu_seeds(cloud)
np.save(os.path.join(ROOT, "gpurun_out", "seeds_c.npy"), np.rint(s * 1e6).astype(np.int32))
print(s.shape)
