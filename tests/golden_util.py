"""Loader for tests/golden/loss_*.npz (written by oracle/make_golden.py from the real reference)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(
        os.path.basename(p)[len("loss_"):-len(".npz")]
        for p in glob.glob(os.path.join(GOLDEN, "loss_*.npz"))
    )


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"loss_{name}.npz"))
    world = int(z["world"])
    flags = {kv.split("=")[0]: bool(int(kv.split("=")[1])) for kv in z["flags"].tolist()}
    pairs = []
    for s in z["pairs"].tolist():
        a, b, w = s.split("|")
        pairs.append((a, b, float(w)))
    emb, ids, out = [], [], []
    for r in range(world):
        e, i, o = {}, {}, {}
        for k in z.files:
            if k.startswith(f"r{r}_emb_"):
                e[k[len(f"r{r}_emb_"):]] = z[k]
            elif k.startswith(f"r{r}_ids_"):
                i[k[len(f"r{r}_ids_"):]] = z[k]
            elif k.startswith(f"r{r}_out_"):
                o[k[len(f"r{r}_out_"):]] = z[k]
        emb.append(e)
        ids.append(i)
        out.append(o)
    return dict(world=world, scale=float(z["scale"]), flags=flags, pairs=pairs, emb=emb, ids=ids,
                out=out)
