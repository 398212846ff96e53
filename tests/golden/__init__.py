import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = sorted(f[:-4] for f in os.listdir(HERE) if f.endswith(".npz") and not f.startswith("pairs_"))   # pairs_*: loader fixtures


def load(name):
    d = dict(np.load(os.path.join(HERE, name + ".npz"), allow_pickle=False))
    d["cfg"] = json.loads(str(d["cfg"]))
    d["U"], d["I"] = int(d["U"]), int(d["I"])
    d["p0"] = {k[3:]: v for k, v in d.items() if k.startswith("p0_")}
    d["p1"] = {k[3:]: v for k, v in d.items() if k.startswith("p1_")}
    return d
