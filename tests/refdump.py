"""Run oracle/_ref/ref_driver (the unmodified reference + shim) and load what it dumps. Test infrastructure."""
import os
import subprocess
import tempfile

import numpy as np

_FILES = {
    "hit_read": "u32", "hit_kmer": "u64", "firstpos_read": "u32", "firstpos_kmer": "u64", "firstpos_pos": "u32",
    "readlen_read": "u32", "readlen_len": "u32", "inv_kmer": "u64", "inv_off": "u64", "inv_read": "u32",
    "conn_x": "u32", "conn_y": "u32", "conn_score": "u64", "comp_off": "u64", "comp_read": "u32", "comp_root": "u32",
    "tree_off": "u64", "tree_x": "u32", "tree_y": "u32",
    "core_id": "u32", "core_kmer_off": "u64", "core_kmer": "u64", "core_read_off": "u64", "core_read": "u32", "purged_off": "u64", "purged_read": "u32",
    "econn_x": "u32", "econn_y": "u32", "econn_score": "u64", "final_id": "u32", "final_off": "u64", "final_read": "u32",
    "tconn_x": "u32", "tconn_y": "u32", "tconn_score": "u64", "spectral_off": "u64", "spectral_member": "u32", "spectral_first": "u32",
    "tail_off": "u64", "tail_vertex": "u32", "amp_off": "u64", "amp_vertex": "u32",
}
_DT = {"u32": np.uint32, "u64": np.uint64}


def run_ref(driver, read_paths, kmer_path, fraction=0.15, min_size=30, min_score=1, threads=1, dump=True, stop_after=0, outdir=None, enrich=0, sc_score=0, full=False, max_size=-1, force_spectral=False):
    tmp = None
    if outdir is None:
        tmp = tempfile.TemporaryDirectory()
        outdir = tmp.name
    cmd = [driver, "run", "--kmers", kmer_path, "--out", outdir, "--threads", str(threads), "--fraction", repr(fraction),
           "--min-size", str(min_size), "--min-score", str(min_score)]
    if not dump:
        cmd.append("--no-dump")
    if stop_after:
        cmd += ["--stop-after", str(stop_after)]
    if enrich:
        cmd += ["--enrich", str(enrich)]
    if sc_score:
        cmd += ["--sc-score", str(sc_score)]
    if full:
        cmd.append("--full")
    if force_spectral:
        cmd.append("--force-spectral")
    if max_size != -1:
        cmd += ["--max-size", str(max_size)]
    cmd += list(read_paths)
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    out = {}
    with open(os.path.join(outdir, "meta.txt")) as f:
        for line in f:
            k, v = line.strip().split("=")
            out[k] = float(v) if "." in v else int(v)
    if dump:
        for name, t in _FILES.items():
            p = os.path.join(outdir, f"{name}.{t}")
            if os.path.exists(p):
                out[name] = np.fromfile(p, dtype=_DT[t])
    if tmp:
        tmp.cleanup()
    return out


def ref_records(driver, read_paths):
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([driver, "records", d] + list(read_paths), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if r.returncode != 0:
            return r.returncode, None, None
        metas, recs = [], []
        with open(os.path.join(d, "records.txt"), newline="\n") as f:
            for line in f:
                line = line[:-1] if line.endswith("\n") else line
                if line.startswith("#META ") or line.startswith("#AGG "):
                    metas.append(line.split(" "))
                else:
                    i, h, s, q = line.split("\t")
                    recs.append((int(i), h, s, q))
        return 0, metas, recs


def ref_kmeriter(driver, seq, k):
    r = subprocess.run([driver, "kmeriter", str(k), seq], check=True, capture_output=True, text=True)
    pos, km = [], []
    for line in r.stdout.split("\n"):
        if line:
            a, b = line.split()
            pos.append(int(a)); km.append(int(b))
    return np.array(km, dtype=np.uint64), np.array(pos, dtype=np.uint32)


def ref_canon(driver, kmer_path):
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([driver, "canon", kmer_path, d], check=True, capture_output=True, text=True)
        kv = dict(l.split("=") for l in r.stdout.split())
        return np.fromfile(os.path.join(d, "canon_kmers.u64"), dtype=np.uint64), int(kv["k"])
