import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["kat2", "config1_mini", "longreads_k15", "longreads_k21", "exceptions_crlf"]
ENRICH_CASES = ["enrich_short", "enrich_long"]     # ref_driver --enrich
FULL_CASES = ["full_short", "full_long"]           # ref_driver --enrich --full (tail / spectral block included)


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    ref.update({k[5:]: int(z[k]) for k in z.files if k.startswith("meta_")})
    return dict(bases=z["bases"].tobytes(), seq_off=z["seq_off"], kmers=z["kmers"], k=int(z["k"]), fraction=float(z["fraction"]),
                min_size=int(z["min_size"]), enrich=int(z["enrich"]) if "enrich" in z.files else 0, ref=ref)


def parse_records(path):
    metas, recs = [], []
    with open(path, newline="\n") as f:
        for line in f:
            line = line[:-1] if line.endswith("\n") else line
            if line.startswith("#META ") or line.startswith("#AGG "):
                metas.append(line.split(" "))
            else:
                i, h, s, q = line.split("\t")
                recs.append((int(i), h, s, q))
    return metas, recs
