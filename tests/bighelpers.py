"""Replay of the full-size fixtures (tests/golden/make_golden_big.py): seeded regeneration of the inputs,
comparison of large tensors through their strided samples + fp64 checksums."""
import json
import os

import numpy as np
import torch

from tests.golden import bigfix
from tests.helpers import GOLDEN_DIR, rel_err


def load_big(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    case = meta["case"]
    data, delta0 = bigfix.regen_data(case), bigfix.regen_delta(case)
    # the regenerated inputs must be the very tensors the reference ran on
    for t, key in ((data, "data__ck"), (delta0, "delta0__ck")):
        ck = bigfix.checksum(t)
        assert torch.allclose(ck, arrays[key], rtol=1e-12, atol=0), \
            "seeded regeneration of %s differs from the fixture (RNG stream changed?)" % key
    floors = {}
    fp = os.path.join(GOLDEN_DIR, name + ".floors.json")
    if os.path.exists(fp):
        with open(fp) as f:
            floors = json.load(f)
    return meta, arrays, data, delta0, floors


def start_params(case, z, delta0, step=0):
    """Start parameters of a step as stored in full (None where only samples exist)."""
    out = []
    for i, n in enumerate(case["chain"]):
        key = ("p0_%d" % i) if step == 0 else ("s%d_param_%d" % (step, i))
        if step == 0 and n == "noise":
            out.append(delta0.clone())
        else:
            out.append(z[key].clone() if key in z else None)
    return out


def big_err(t, z, key, case):
    """rel. error of `t` against fixture entry `key`: full tensor when stored in full, else the strided
    sample; second value = checksum deviation of the whole tensor (0.0 when stored in full)."""
    if key in z:
        return rel_err(t, z[key]), 0.0
    e = rel_err(bigfix.strided(t, case), z[key + "__s"])
    return e, bigfix.checksum_err(t, z[key + "__ck"])
