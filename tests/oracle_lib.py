"""ctypes binding of oracle/libhga_oracle.so — TEST INFRASTRUCTURE ONLY (never imported by the product)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)


class OrcReads(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("seq_off", u64p), ("seq", C.c_void_p), ("hdr_off", u64p), ("hdr", C.c_void_p),
                ("qual_off", u64p), ("qual", C.c_void_p), ("file_index", i32p), ("n_files", C.c_int),
                ("f_records", u64p), ("f_min", u64p), ("f_max", u64p), ("f_avg", u64p), ("f_total", u64p), ("f_type", i32p),
                ("a_records", C.c_uint64), ("a_min", C.c_uint64), ("a_max", C.c_uint64), ("a_avg", C.c_uint64), ("a_total", C.c_uint64)]


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.orc_kmer_windows.restype = C.c_uint64
        lib.orc_free.argtypes = [C.c_void_p]

    # -- KmerIterator
    def kmer_windows(self, seq: bytes, k: int):
        n = max(0, len(seq) - k + 1) if k <= 32 else 0
        km = np.zeros(n + 1, dtype=np.uint64)
        pos = np.zeros(n + 1, dtype=np.uint32)
        got = self.lib.orc_kmer_windows(C.c_char_p(seq), C.c_uint64(len(seq)), C.c_int(k), km.ctypes.data_as(u64p), pos.ctypes.data_as(u32p))
        return km[:got], pos[:got]

    def load_kmers(self, path):
        out = u64p(); n = C.c_uint64(); k = C.c_int()
        rc = self.lib.orc_load_kmers(path.encode(), C.byref(out), C.byref(n), C.byref(k))
        if rc:
            raise RuntimeError(f"orc_load_kmers rc={rc}")
        arr = _np(out, n.value, np.uint64)
        self.lib.orc_free(out)
        return arr, k.value

    def load_reads(self, paths):
        r = OrcReads()
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        rc = self.lib.orc_load_reads(arr, C.c_int(len(paths)), C.byref(r))
        if rc:
            return rc, None
        n = r.n_reads
        so = _np(r.seq_off, n + 1, np.uint64); ho = _np(r.hdr_off, n + 1, np.uint64); qo = _np(r.qual_off, n + 1, np.uint64)
        d = dict(n_reads=n, seq_off=so, hdr_off=ho, qual_off=qo,
                 seq=C.string_at(r.seq, int(so[-1])), hdr=C.string_at(r.hdr, int(ho[-1])), qual=C.string_at(r.qual, int(qo[-1])),
                 file_index=_np(r.file_index, n, np.int32),
                 f_records=_np(r.f_records, r.n_files, np.uint64), f_min=_np(r.f_min, r.n_files, np.uint64),
                 f_max=_np(r.f_max, r.n_files, np.uint64), f_avg=_np(r.f_avg, r.n_files, np.uint64),
                 f_total=_np(r.f_total, r.n_files, np.uint64), f_type=_np(r.f_type, r.n_files, np.int32),
                 a_records=r.a_records, a_min=r.a_min, a_max=r.a_max, a_avg=r.a_avg, a_total=r.a_total)
        self.lib.orc_free_reads(C.byref(r))
        return 0, d

    def scan(self, seq: bytes, seq_off, k, kmers_sorted):
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        kmers_sorted = np.ascontiguousarray(kmers_sorted, dtype=np.uint64)
        n_reads = seq_off.shape[0] - 1
        ro = u64p(); kid = u32p(); pos = u32p()
        rc = self.lib.orc_scan(C.c_char_p(seq), seq_off.ctypes.data_as(u64p), C.c_uint64(n_reads), C.c_int(k),
                               kmers_sorted.ctypes.data_as(u64p), C.c_uint64(kmers_sorted.shape[0]), C.byref(ro), C.byref(kid), C.byref(pos))
        if rc:
            raise RuntimeError(f"orc_scan rc={rc}")
        row_off = _np(ro, n_reads + 1, np.uint64)
        E = int(row_off[-1])
        out = (row_off, _np(kid, E, np.uint32), _np(pos, E, np.uint32))
        for p in (ro, kid, pos):
            self.lib.orc_free(p)
        return out

    def index(self, row_off, hit_kid, n_kmers):
        row_off = np.ascontiguousarray(row_off, dtype=np.uint64); hit_kid = np.ascontiguousarray(hit_kid, dtype=np.uint32)
        n_reads = row_off.shape[0] - 1
        off = u64p(); rd = u32p()
        self.lib.orc_index(row_off.ctypes.data_as(u64p), hit_kid.ctypes.data_as(u32p), C.c_uint64(n_reads), C.c_uint64(n_kmers), C.byref(off), C.byref(rd))
        inv_off = _np(off, n_kmers + 1, np.uint64)
        inv_read = _np(rd, int(inv_off[-1]), np.uint32)
        self.lib.orc_free(off); self.lib.orc_free(rd)
        return inv_off, inv_read

    def connections(self, row_off, hit_kid, inv_off, inv_read, min_score=1, pivots=None):
        row_off = np.ascontiguousarray(row_off, dtype=np.uint64); hit_kid = np.ascontiguousarray(hit_kid, dtype=np.uint32)
        inv_off = np.ascontiguousarray(inv_off, dtype=np.uint64); inv_read = np.ascontiguousarray(inv_read, dtype=np.uint32)
        n_reads = row_off.shape[0] - 1
        n = C.c_uint64(); cx = u32p(); cy = u32p(); cs = u64p()
        if pivots is None:
            pp, npv = None, 0
        else:
            pivots = np.ascontiguousarray(pivots, dtype=np.uint32)
            pp, npv = pivots.ctypes.data_as(u32p), pivots.shape[0]
        self.lib.orc_connections(row_off.ctypes.data_as(u64p), hit_kid.ctypes.data_as(u32p), C.c_uint64(n_reads), inv_off.ctypes.data_as(u64p),
                                 inv_read.ctypes.data_as(u32p), pp, C.c_uint64(npv), C.c_uint64(min_score), C.byref(n), C.byref(cx), C.byref(cy), C.byref(cs))
        out = (_np(cx, n.value, np.uint32), _np(cy, n.value, np.uint32), _np(cs, n.value, np.uint64))
        for p in (cx, cy, cs):
            self.lib.orc_free(p)
        return out

    def canonical_sort(self, cx, cy, cs):
        cx = np.array(cx, dtype=np.uint32); cy = np.array(cy, dtype=np.uint32); cs = np.array(cs, dtype=np.uint64)
        self.lib.orc_canonical_sort(C.c_uint64(cx.shape[0]), cx.ctypes.data_as(u32p), cy.ctypes.data_as(u32p), cs.ctypes.data_as(u64p))
        return cx, cy, cs

    def union_find(self, ex, ey, min_size=30, max_size=-1, restricted=None):
        ex = np.ascontiguousarray(ex, dtype=np.uint32); ey = np.ascontiguousarray(ey, dtype=np.uint32)
        nc = C.c_uint64(); co = u64p(); cm = u32p(); to = u64p(); tx = u32p(); ty = u32p()
        if restricted is None:
            rp, nr = None, 0
        else:
            restricted = np.ascontiguousarray(restricted, dtype=np.uint32)
            rp, nr = restricted.ctypes.data_as(u32p), restricted.shape[0]
        self.lib.orc_union_find(C.c_uint64(ex.shape[0]), ex.ctypes.data_as(u32p), ey.ctypes.data_as(u32p), rp, C.c_uint64(nr), C.c_int(min_size),
                                C.c_int(max_size), C.byref(nc), C.byref(co), C.byref(cm), C.byref(to), C.byref(tx), C.byref(ty))
        comp_off = _np(co, nc.value + 1, np.uint64); tree_off = _np(to, nc.value + 1, np.uint64)
        out = (comp_off, _np(cm, int(comp_off[-1]), np.uint32), tree_off, _np(tx, int(tree_off[-1]), np.uint32), _np(ty, int(tree_off[-1]), np.uint32))
        for p in (co, cm, to, tx, ty):
            self.lib.orc_free(p)
        return out

    # -- whole hot path on in-memory reads: returns a dict mirroring the ref_driver dump
    def run(self, seq: bytes, seq_off, k, kmers_sorted, fraction=0.15, min_size=30, min_score=1, sc_score=0, max_size=-1):
        row_off, kid, pos = self.scan(seq, seq_off, k, kmers_sorted)
        inv_off, inv_read = self.index(row_off, kid, len(kmers_sorted))
        if sc_score > 0:
            # --sc_score S (run_clustering :749-752): pivots = components with >= S discriminative k-mers, keep score > S
            pivots = (np.nonzero(np.diff(row_off.astype(np.int64)) >= sc_score)[0] + 1).astype(np.uint32)
            cx, cy, cs = self.connections(row_off, kid, inv_off, inv_read, min_score=sc_score, pivots=pivots)
            sx, sy, ss = self.canonical_sort(cx, cy, cs)
            n = int((ss > sc_score).sum())
            cut = sc_score
        else:
            cx, cy, cs = self.connections(row_off, kid, inv_off, inv_read, min_score=min_score)
            sx, sy, ss = self.canonical_sort(cx, cy, cs)
            n = int(len(sx) * fraction)
            cut = int(ss[n - 1]) if n > 0 else 0
        comp = self.union_find(sx[:n], sy[:n], min_size=min_size, max_size=max_size)
        return dict(row_off=row_off, hit_kid=kid, hit_pos=pos, inv_off=inv_off, inv_read=inv_read, conn=(sx, sy, ss), cut_n=n, cut_score=cut,
                    comp=comp)


class Engine:
    """Mutable engine state after the scaffold stage: merge_components / get_connections / get_component_ids."""

    def __init__(self, orc, row_off, hit_kid, n_kmers, inv_off, inv_read):
        self.lib = orc.lib
        self.lib.orc_engine_new.restype = C.c_void_p
        for f in ("orc_engine_ids", "orc_engine_component_kmers", "orc_engine_component_reads", "orc_engine_index_list"):
            getattr(self.lib, f).restype = C.c_uint64
        row_off = np.ascontiguousarray(row_off, dtype=np.uint64); hit_kid = np.ascontiguousarray(hit_kid, dtype=np.uint32)
        inv_off = np.ascontiguousarray(inv_off, dtype=np.uint64); inv_read = np.ascontiguousarray(inv_read, dtype=np.uint32)
        self.n_reads = row_off.shape[0] - 1
        self.n_kmers = n_kmers
        self.h = C.c_void_p(self.lib.orc_engine_new(row_off.ctypes.data_as(u64p), hit_kid.ctypes.data_as(u32p), C.c_uint64(self.n_reads), C.c_uint64(n_kmers),
                                                    inv_off.ctypes.data_as(u64p), inv_read.ctypes.data_as(u32p)))

    def close(self):
        if self.h:
            self.lib.orc_engine_free(self.h)
            self.h = None

    def merge(self, comp_off, comp_member):
        comp_off = np.ascontiguousarray(comp_off, dtype=np.uint64); comp_member = np.ascontiguousarray(comp_member, dtype=np.uint32)
        n = comp_off.shape[0] - 1
        ids = np.zeros(max(n, 1), dtype=np.uint32)
        self.lib.orc_engine_merge(self.h, C.c_uint64(n), comp_off.ctypes.data_as(u64p), comp_member.ctypes.data_as(u32p), ids.ctypes.data_as(u32p))
        return ids[:n]

    def ids(self, min_size):
        out = np.zeros(self.n_reads + 1, dtype=np.uint32)
        n = self.lib.orc_engine_ids(self.h, C.c_uint64(min_size), out.ctypes.data_as(u32p))
        return out[:n].copy()

    def _list(self, fn, key, ktype):
        n = getattr(self.lib, fn)(self.h, ktype(key), None)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        getattr(self.lib, fn)(self.h, ktype(key), out.ctypes.data_as(u32p))
        return out[:n]

    def component_kmers(self, cid):
        return self._list("orc_engine_component_kmers", int(cid), C.c_uint32)

    def component_reads(self, cid):
        return self._list("orc_engine_component_reads", int(cid), C.c_uint32)

    def index_list(self, kmer_id):
        return self._list("orc_engine_index_list", int(kmer_id), C.c_uint64)

    def index(self):
        lists = [self.index_list(k) for k in range(self.n_kmers)]
        off = np.zeros(self.n_kmers + 1, dtype=np.uint64)
        np.cumsum([len(l) for l in lists], out=off[1:])
        return off, (np.concatenate(lists) if lists else np.zeros(0, dtype=np.uint32)).astype(np.uint32)

    def connections(self, pivots, min_score):
        pivots = np.ascontiguousarray(pivots, dtype=np.uint32)
        n = C.c_uint64(); cx = u32p(); cy = u32p(); cs = u64p()
        self.lib.orc_engine_connections(self.h, pivots.ctypes.data_as(u32p), C.c_uint64(pivots.shape[0]), C.c_uint64(min_score), C.byref(n), C.byref(cx),
                                        C.byref(cy), C.byref(cs))
        out = (_np(cx, n.value, np.uint32), _np(cy, n.value, np.uint32), _np(cs, n.value, np.uint64))
        for p in (cx, cy, cs):
            self.lib.orc_free(p)
        return out


def enrich(orc, res, n_kmers, min_size=30, enrich_min=20):
    """run_clustering :764 and :785-794 on the result of Oracle.run (tail / spectral block skipped: the reference's own path
    when it has at most two scaffold components or no strong tail connection). Returns a dict mirroring ref_driver --enrich."""
    co, cm = res["comp"][0], res["comp"][1]
    eng = Engine(orc, res["row_off"], res["hit_kid"], n_kmers, res["inv_off"], res["inv_read"])
    try:
        eng.merge(co, cm)
        cores = eng.ids(min_size)
        core_kmers = [eng.component_kmers(c) for c in cores]
        core_reads = [np.sort(eng.component_reads(c)) for c in cores]
        purged_off, purged_read = eng.index()
        ex, ey, es = orc.canonical_sort(*eng.connections(cores, enrich_min))
        eo, em, _, _, _ = orc.union_find(ex, ey, min_size=2, max_size=-1, restricted=cores)
        eng.merge(eo, em)
        final = eng.ids(min_size)
        final_reads = [np.sort(eng.component_reads(c)) for c in final]
    finally:
        eng.close()
    order = np.argsort([int(r[0]) for r in final_reads], kind="stable") if len(final) else []
    return dict(core_id=cores, core_kmers=core_kmers, core_reads=core_reads, purged_off=purged_off, purged_read=purged_read,
                econn=(ex, ey, es), final_id=np.array([final[i] for i in order], dtype=np.uint32), final_reads=[final_reads[i] for i in order])


_cached = None


def load():
    global _cached
    if _cached is None:
        so = os.path.join(ODIR, "libhga_oracle.so")
        src = os.path.join(ODIR, "hga_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-C", ODIR, "port"], check=True, capture_output=True)
        _cached = Oracle(C.CDLL(so))
    return _cached
