import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, horayzon_b200 as hb
from horayzon_b200 import resident
for n in (1025, 1201):
    x, y, z = hb.synthetic.sinusoid_dem(n, n, 90.0, 1200.0, 40000.0, 2, 6)
    sc = resident.Scene(hb.synthetic.rearrange_pad_buffer(x, y, z), n, n); st = sc.stats(); sc.close()
    print("n=%d morton=%s prims=%d nodes4=%d prims/nodes=%.3f children/node=%.3f" % (n, os.environ.get("HZB_MORTON", "default"), st["num_prims"], st["num_nodes"], st["num_prims"] / st["num_nodes"], (st["num_nodes"] - 1 + st["num_prims"]) / st["num_nodes"]))
