"""Host-side mirror of the reference's operator interface for the categorization hot path.

Names, argument meaning and error behaviour follow the reference (paths relative to /root/reference/src):
  load_text_file_kmers      read_clustering.cpp:18-33
  SequenceRecords           common/SequenceRecordIterator.{h,cpp}  (record rules, 1-based global ReadID, MetaData)
  ReadClusteringConfig      clustering/ReadClusteringEngine.h:138-148
  ReadClusteringEngine      clustering/ReadClusteringEngine.h:150-199: construct_indices, get_connections,
                            get_all_connections, union_find, run_clustering (:699-802) without its tail / spectral block
                            (:768-777), export_components (:804-826)

All arithmetic on reads happens in libhga_b200.so on the GPU; this module only parses files, owns the handle and
shapes results. It never touches oracle/.
"""
from dataclasses import dataclass
import os

import numpy as np

from . import capi

_CODE = np.zeros(256, dtype=np.uint64)
_COMP = np.zeros(256, dtype=np.uint64)
for _c, _v in zip(b"ACGT", (0, 1, 2, 3)):
    _CODE[_c] = _v
    _COMP[_c] = 3 - _v
# any other byte: 0 in BOTH tables (unordered_map::operator[] default-inserts, KmerIterator.cpp:56,62)


def canonical_kmer(seq: bytes, k: int) -> int:
    """First window of KmerIterator(seq, k) — what load_text_file_kmers stores for a line."""
    if k > 32:
        raise ValueError("Kmer size is too big")            # KmerIterator.cpp:24-26
    if k == 0 or len(seq) < k:
        return 0
    b = np.frombuffer(seq[:k], dtype=np.uint8)
    sh = np.arange(k, dtype=np.uint64)
    fwd = int(np.bitwise_or.reduce(_CODE[b] << (np.uint64(2) * (np.uint64(k - 1) - sh))))
    rev = int(np.bitwise_or.reduce(_COMP[b] << (np.uint64(2) * sh)))
    return min(fwd, rev)


def load_text_file_kmers(path):
    """-> (sorted unique canonical k-mers as uint64 array, k). k is the length of the LAST line (.cpp:27,32)."""
    k = 0
    vals = []
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError:
        return np.zeros(0, dtype=np.uint64), 0                # ifstream on a missing file: empty set
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()                                           # std::getline yields no extra line after a final '\n'
    if lines:
        lens = {len(l) for l in lines}
        if len(lens) == 1 and (k := lens.pop()) and 0 < k <= 32:
            # fast path: all lines the same length -> vectorised
            arr = np.frombuffer(b"".join(lines), dtype=np.uint8).reshape(len(lines), k)
            sh = np.arange(k, dtype=np.uint64)
            fwd = np.bitwise_or.reduce(_CODE[arr] << (np.uint64(2) * (np.uint64(k - 1) - sh)), axis=1)
            rev = np.bitwise_or.reduce(_COMP[arr] << (np.uint64(2) * sh), axis=1)
            return np.unique(np.minimum(fwd, rev)), k
        for l in lines:
            k = len(l)
            vals.append(canonical_kmer(l, k))
    return np.unique(np.array(vals, dtype=np.uint64)), k


@dataclass
class MetaData:
    filename: str = ""
    records: int = 0
    min_read_length: int = 2 ** 64 - 1
    max_read_length: int = 0
    avg_read_length: int = 0
    total_bases: int = 0
    file_type: str = "UNKNOWN"

    def repr(self):
        return (f"{self.filename}:\n- {self.records} reads\n- {self.total_bases} total bases\n- {self.avg_read_length} average read length\n"
                f"- {self.max_read_length} max read length\n- {self.min_read_length} min read length\n\n")


class SequenceRecords:
    """All records of the given FASTA/FASTQ files, read in one pass with the reference's rules:
    FASTQ = 4 lines/record, FASTA = 2 lines/record (multi-line FASTA unsupported), format sniffed per file from its
    first record, the line stream continues across file boundaries, header = line minus its first character,
    ReadID = 1, 2, ... across files. Errors the reference raises as C++ exceptions are raised here as
    ValueError (invalid_argument / logic_error) or IndexError (out_of_range)."""

    def __init__(self, paths, annotate=False):
        self.paths = list(paths)
        self.annotate = annotate
        self.file_meta = [MetaData() for _ in self.paths]
        self.meta = MetaData()
        self._load()
        self.categories = len(self.paths) if annotate else 1

    def _open(self, pos):
        try:
            with open(self.paths[pos], "rb") as f:
                data = f.read()
        except OSError:
            raise ValueError(f'File with path "{self.paths[pos]}" does not exist')
        lines = data.split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        # sniff (load_file_at_position :85-99)
        if len(lines) < 2:
            raise ValueError("File is empty")
        h0 = lines[0][:1]
        if h0 == b"@":
            if len(lines) < 3:
                raise ValueError("File is empty")
            if lines[2][:1] == b"+":
                self._method = 4
                self._type = "FASTQ"
        elif h0 == b">":
            self._method = 2
            self._type = "FASTA"
        else:
            raise ValueError("Unrecognized file format")
        return lines

    def _load(self):
        self._method, self._type = 4, "FASTQ"
        cur = 0
        lines = self._open(0)
        at = 0
        seqs, hdrs, quals, fidx = [], [], [], []
        filenames = []
        prev_file = -1

        def next_line():
            nonlocal cur, lines, at
            while at >= len(lines):
                if cur + 1 >= len(self.paths):
                    return None
                cur += 1
                lines = self._open(cur)
                at = 0
            at += 1
            return lines[at - 1]

        while True:
            n = self._method
            rec = []
            for _ in range(n):
                l = next_line()
                if l is None:
                    break
                rec.append(l)
            if len(rec) < n:
                break
            if len(rec[0]) == 0:
                raise IndexError("basic_string::substr")     # header.substr(1) on an empty header
            hdrs.append(rec[0][1:]); seqs.append(rec[1]); quals.append(rec[3] if n == 4 else b""); fidx.append(cur)
            if cur != prev_file:
                self.file_meta[cur] = MetaData(filename=os.path.basename(self.paths[cur]), file_type=self._type)
                filenames.append(self.file_meta[cur].filename)
                prev_file = cur
            fm, L = self.file_meta[cur], len(rec[1])
            fm.total_bases += L; fm.min_read_length = min(fm.min_read_length, L); fm.max_read_length = max(fm.max_read_length, L)
            fm.records += 1; fm.avg_read_length += L
            m = self.meta
            m.total_bases += L; m.min_read_length = min(m.min_read_length, L); m.records += 1; m.avg_read_length += L
        if self.meta.records == 0:
            raise ValueError("File is empty")
        self.meta.avg_read_length //= self.meta.records       # aggregate max_read_length stays 0 (never updated, :59-62)
        self.meta.filename = "__".join(filenames)
        for fm in self.file_meta:
            if not fm.records:
                # data.avg_read_length /= data.records for every file (:69-71): SIGFPE in the reference when a file got no record
                raise ZeroDivisionError("a read file contributed no record (division by zero in load_meta_data)")
            fm.avg_read_length //= fm.records
        self.headers, self.qualities = hdrs, quals
        self.file_index = np.array(fidx, dtype=np.int32)
        lens = np.fromiter((len(s) for s in seqs), dtype=np.uint64, count=len(seqs))
        self.seq_off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum(lens, out=self.seq_off[1:])
        self.bases = b"".join(seqs)

    @property
    def n_reads(self):
        return len(self.headers)

    def sequence(self, read_id):
        i = read_id - 1
        return self.bases[int(self.seq_off[i]):int(self.seq_off[i + 1])]

    def fastx_string(self, read_id):
        i = read_id - 1
        if self.qualities[i]:
            return b"@" + self.headers[i] + b"\n" + self.sequence(read_id) + b"\n+\n" + self.qualities[i]
        return b">" + self.headers[i] + b"\n" + self.sequence(read_id)


@dataclass
class ReadClusteringConfig:
    scaffold_component_min_size: int = 30
    scaffold_component_max_size: int = -1
    scaffold_forming_fraction: float = 0.15
    scaffold_forming_score: int = 0
    enrichment_connections_min_score: int = 20
    tail_amplification_min_score: int = 40
    threads: int = 1
    spectral_dims: int = 16
    force_spectral: bool = False


class ReadClusteringEngine:
    """GPU-backed stand-in for the reference engine's hot-path stages. Component ids are read ids, as in the
    reference before any merge (ReadComponent ctor, ReadClusteringEngine.h:38-43)."""

    def __init__(self, reader: SequenceRecords, config: ReadClusteringConfig = None, device=0, dist=None):
        """dist: an initialised torch.distributed module (one rank per GPU) or None. With it every rank holds the SAME reader and k-mer set; the hot path
        (construct_indices ... union_find) runs on this rank's contiguous shard of the reads with the exchanges of hga_comm.cu, hga_comm_gather_root then
        leaves a complete single-GPU state on rank 0, and the rest of run_clustering runs there (what `categorization --gpus N` does with rank threads)."""
        self.reader = reader
        self.config = config or ReadClusteringConfig()
        self.device = device
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.handle = None
        self.kmers = None

    def close(self):
        if self.handle is not None:
            self.handle.close()
            self.handle = None

    # construct_indices(discriminative_kmers, k)  — .cpp:234-299
    def construct_indices(self, discriminative_kmers, k):
        self.close()
        self.kmers = np.unique(np.asarray(discriminative_kmers, dtype=np.uint64))
        self.handle = capi.Handle(self.kmers, k, device=self.device)
        if self.world > 1:
            from . import parallel
            off = np.asarray(self.reader.seq_off, dtype=np.uint64)
            bounds = parallel.shard_bounds(np.diff(off.astype(np.int64)), self.world)
            lo, hi = bounds[self.rank], bounds[self.rank + 1]
            uid = parallel.broadcast_unique_id(self.dist, self.rank, capi.comm_unique_id)
            self.handle.comm_init(uid, self.rank, self.world, self.reader.n_reads)
            bases = self.reader.bases[int(off[lo]):int(off[hi])]
            self.handle.scan(bases, off[lo:hi + 1] - off[lo], read_id_base=lo + 1)
        else:
            self.handle.scan(self.reader.bases, self.reader.seq_off, read_id_base=1)
        self.handle.build_index()
        return 0

    def component_ids(self):
        """ids of the reads that have at least one discriminative k-mer (= keys of component_index)"""
        row_off, _, _ = self.handle.get_hits()
        return (np.nonzero(np.diff(row_off.astype(np.int64)) > 0)[0] + 1).astype(np.uint32)

    # get_connections(component_ids, min_score) — .cpp:301-333; returns the DIRECTED list sorted by the canonical
    # order (score desc, min id asc, max id asc, x asc); x is always one of component_ids.
    def get_connections(self, component_ids, min_score):
        pivots = np.asarray(component_ids, dtype=np.uint32)
        self.handle.pair_count(min_score=min_score, pivots=pivots)
        x, y, s, _ = self.handle.get_pairs()
        is_p = np.zeros(self.reader.n_reads + 2, dtype=bool)
        is_p[pivots] = True
        fx = np.concatenate([x[is_p[x]], y[is_p[y]]]); fy = np.concatenate([y[is_p[x]], x[is_p[y]]]); fs = np.concatenate([s[is_p[x]], s[is_p[y]]])
        return _canonical_sort(fx, fy, fs.astype(np.uint64))

    # get_all_connections(min_score) — .cpp:335-339
    def get_all_connections(self, min_score):
        self.handle.pair_count(min_score=min_score)
        x, y, s, _ = self.handle.get_pairs()
        return _canonical_sort(np.concatenate([x, y]), np.concatenate([y, x]), np.concatenate([s, s]).astype(np.uint64))

    # the hot first third of run_clustering (.cpp:737-765): returns the scaffold components (lists of read ids)
    def _select_scaffold_edges(self):
        cfg = self.config
        if cfg.scaffold_forming_score > 0:
            row_off, _, _ = self.handle.get_hits()
            ids = (np.nonzero(np.diff(row_off.astype(np.int64)) >= cfg.scaffold_forming_score)[0] + 1).astype(np.uint32)
            self.handle.pair_count(min_score=cfg.scaffold_forming_score, pivots=ids)
            self.handle.select_edges(score_threshold=cfg.scaffold_forming_score)
        else:
            self.handle.pair_count(min_score=1)
            self.handle.select_edges(fraction=cfg.scaffold_forming_fraction)

    def scaffold_components(self):
        cfg = self.config
        self._select_scaffold_edges()
        if cfg.scaffold_component_max_size != -1:
            raise NotImplementedError("--sc_max_size: the size-limited components come from run_clustering (hga_enrich_ex), not from the GPU components stage")
        self.handle.components(min_size=cfg.scaffold_component_min_size)
        comp = self.handle.get_components()
        label = comp["label"]
        out = []
        for root in comp["comp_label"]:
            out.append((np.nonzero(label == root)[0] + comp["read_id_first"]).astype(np.uint32))
        return out


    # run_clustering(discriminative_kmers, k) — .cpp:699-802. Returns the final component ids; their members are in
    # self.final_components (id -> ascending read ids). The tail / spectral merge of scaffold components (:768-777) is part of the
    # run (hga_enrich_full); with tail_block=False every scaffold component becomes a core, which is the reference's own path when it
    # finds no strong tail connection.
    def run_clustering(self, discriminative_kmers, k, tail_block=True):
        cfg = self.config
        self.construct_indices(discriminative_kmers, k)
        if self.world > 1 and cfg.force_spectral:
            raise NotImplementedError("--spectral needs every pair on one GPU: single-GPU only")
        if cfg.force_spectral:
            # :739-746: get_all_connections(5) on the GPU, spectral clustering of the whole data set on the host (an S x S
            # eigen-problem over all connected reads, as in the reference: small inputs only), merge_components, ids with >= min size
            cx, cy, cs = self.get_all_connections(5)
            self.final_components = forced_spectral_components(cx, cy, cs, cfg.spectral_dims, cfg.scaffold_component_min_size)
            self.assignment = np.zeros(self.reader.n_reads, dtype=np.uint32)
            for fid, members in self.final_components.items():
                self.assignment[members - 1] = fid
            return list(self.final_components)
        if self.world > 1 and cfg.scaffold_forming_score > 0:
            raise NotImplementedError("--sc_score (a pivot subset) is single-GPU only")
        self._select_scaffold_edges()
        if self.world > 1:
            # the stages after the scaffold union_find run on rank 0 (hga_comm_gather_root); every rank gets the result
            self.handle.components(min_size=cfg.scaffold_component_min_size)
            self.handle.comm_gather_root()
            box = [None]
            if self.rank == 0:
                self._merge_and_enrich(tail_block)
                box = [(self.final_components, self.assignment)]
            self.dist.broadcast_object_list(box, src=0)
            self.final_components, self.assignment = box[0]
            return [int(v) for v in self.final_components]
        return self._merge_and_enrich(tail_block)

    def _merge_and_enrich(self, tail_block):
        cfg = self.config
        if tail_block:
            self.handle.enrich_full(self.reader.seq_off, min_size=cfg.scaffold_component_min_size, enrichment_min_score=cfg.enrichment_connections_min_score,
                                    max_size=cfg.scaffold_component_max_size, tail_amplification_min_score=cfg.tail_amplification_min_score,
                                    spectral_dims=cfg.spectral_dims)
        else:
            self.handle.enrich(min_size=cfg.scaffold_component_min_size, enrichment_min_score=cfg.enrichment_connections_min_score,
                               max_size=cfg.scaffold_component_max_size)
        e = self.handle.get_enrichment()
        off = e["final_off"].astype(np.int64)
        self.final_components = {int(fid): e["final_read"][off[i]:off[i + 1]] for i, fid in enumerate(e["final_id"])}
        self.assignment = e["assignment"]
        return [int(v) for v in e["final_id"]]

    # export_components(component_ids, directory_path) — .cpp:804-826
    def export_components(self, component_ids, directory_path):
        import shutil
        shutil.rmtree(directory_path, ignore_errors=True)
        os.makedirs(directory_path)
        wanted = set(int(c) for c in component_ids)
        files = {c: open(os.path.join(directory_path, f"#{c}.fa"), "wb") for c in wanted}
        for rid in range(1, self.reader.n_reads + 1):
            c = int(self.assignment[rid - 1])
            if c in wanted:
                files[c].write(self.reader.fastx_string(rid) + b"\n")
        for f in files.values():
            f.close()
        print(f"Exported {len(wanted)} components")


def forced_spectral_components(conn_x, conn_y, conn_score, dims, min_size):
    """run_clustering with --spectral (.cpp:739-746) after get_all_connections(5): spectral_clustering over the directed connection
    list in canonical order, merge_components (element [0] of a cluster survives), get_component_ids(min_size). Returns
    {surviving id: ascending member read ids}, ordered by the smallest member."""
    clusters = [c for c in capi.spectral_clustering(conn_x, conn_y, conn_score, dims) if len(c) >= max(int(min_size), 1)]
    clusters.sort(key=lambda c: int(c.min()))
    return {int(c[0]): np.sort(c).astype(np.uint32) for c in clusters}


def _canonical_sort(x, y, s):
    lo, hi = np.minimum(x, y), np.maximum(x, y)
    order = np.lexsort((x, hi, lo, np.iinfo(np.uint64).max - s))
    return x[order], y[order], s[order]
