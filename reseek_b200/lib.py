"""ctypes binding of libreseek_b200.so (include/reseek_b200.h).  No compute happens in Python."""
import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("RSK_LIB", HERE / "libreseek_b200.so"))  # RSK_LIB: developer override for A/B builds

NFEAT = 8
TABLE_FLOATS = 2192
MODE_FAST, MODE_SENSITIVE, MODE_VERYSENSITIVE = 1, 2, 3
KEEP_HITS, KEEP_ALL = 0, 1
HIT_MU_REJECTED, HIT_HAS_EVALUE, HIT_REPORTED, HIT_MKF, HIT_GLOBAL = 1, 2, 4, 8, 16
FLT_MAX = float(np.finfo(np.float32).max)


class ReseekB200Error(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("gap_open", C.c_float), ("gap_ext", C.c_float), ("min_fwd_score", C.c_float),
                ("omega", C.c_float), ("omega_fwd", C.c_float), ("mu_gap_open", C.c_int32),
                ("mu_gap_ext", C.c_int32), ("mkfl", C.c_uint32), ("mkf_x1", C.c_int32), ("mkf_x2", C.c_int32),
                ("mkf_min_hsp_score", C.c_int32), ("mkf_min_mega_hsp_score", C.c_float),
                ("max_evalue", C.c_double), ("weights", C.c_float * NFEAT), ("tables", C.c_float * TABLE_FLOATS)]


class ChainsHost(C.Structure):
    _fields_ = [("n", C.c_uint32), ("total", C.c_uint64), ("len", C.c_void_p), ("prof", C.c_void_p),
                ("mu", C.c_void_p), ("xyz", C.c_void_p), ("selfrev", C.c_void_p)]


class CoordsHost(C.Structure):
    _fields_ = [("n", C.c_uint32), ("total", C.c_uint64), ("len", C.c_void_p), ("aa", C.c_void_p), ("xyz", C.c_void_p)]


class HitView(C.Structure):
    _fields_ = [("hit", C.c_void_p), ("path", C.c_char_p), ("label_a", C.c_char_p), ("label_b", C.c_char_p),
                ("seq_a", C.c_char_p), ("seq_b", C.c_char_p), ("len_a", C.c_uint32), ("len_b", C.c_uint32)]


class SearchOpts(C.Structure):
    _fields_ = [("keep", C.c_int32), ("want_paths", C.c_int32), ("skip_evalue", C.c_int32), ("reserved", C.c_int32)]


class PrefilterOpts(C.Structure):
    _fields_ = [("index_mode", C.c_int32), ("rsb_size", C.c_uint32), ("no_kl_swap", C.c_int32), ("raw_only", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("pairs", C.c_uint64), ("mu_filter_in", C.c_uint64), ("mu_filter_rejected", C.c_uint64),
                ("mu_saturated", C.c_uint64), ("sw_pairs", C.c_uint64), ("sw_cells", C.c_uint64),
                ("evalue_pairs", C.c_uint64), ("hits", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("sw_kernel_ms", C.c_float),
                ("mu_kernel_ms", C.c_float), ("lddt_kernel_ms", C.c_float), ("total_ms", C.c_float),
                ("mkf_kernel_ms", C.c_float), ("sw_kernel_launches", C.c_uint32), ("mkf_pairs", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CommStats(C.Structure):
    _fields_ = [("bytes_sent", C.c_uint64), ("bytes_recv", C.c_uint64), ("collective_ms", C.c_float), ("collectives", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


COMM_ID_BYTES = 128

# numpy view of rsk_hit (include/reseek_b200.h)
HIT_DTYPE = np.dtype([
    ("a", np.uint32), ("b", np.uint32), ("score", np.float32), ("lo_a", np.uint32), ("lo_b", np.uint32),
    ("hi_a", np.uint32), ("hi_b", np.uint32), ("ids", np.uint32), ("gaps", np.uint32), ("lddt", np.float32),
    ("ts", np.float32), ("pvalue", np.float32), ("evalue", np.float32), ("qual", np.float32),
    ("mu_score", np.float32), ("mu_fwd", np.int32), ("mu_rev", np.int32), ("flags", np.uint32),
    ("path_len", np.uint32), ("path_off", np.uint64)], align=True)

_lib = None


def load_library():
    """Load libreseek_b200.so; raise (loudly) when it has not been built - there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ReseekB200Error(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C reseek_b200/csrc` (no CPU fallback exists)")
    L = C.CDLL(str(LIB_PATH))
    L.rsk_version.restype = C.c_char_p
    L.rsk_last_error.restype = C.c_char_p
    L.rsk_feature_bgfreq.restype = C.POINTER(C.c_float)
    L.rsk_mu_matrix_i8.restype = C.POINTER(C.c_int8)
    L.rsk_mu_kmer_matrix_i8.restype = C.POINTER(C.c_int8)
    L.rsk_mu_matrix_f32.restype = C.POINTER(C.c_float)
    L.rsk_chainset_count.restype = C.c_uint32
    L.rsk_chainset_residues.restype = C.c_uint64
    L.rsk_results_count.restype = C.c_uint64
    L.rsk_results_hits.restype = C.c_void_p
    L.rsk_results_paths.restype = C.c_void_p
    L.rsk_results_paths_bytes.restype = C.c_uint64
    for fn in (L.rsk_pvalue, L.rsk_evalue, L.rsk_qual):
        fn.restype = C.c_double
        fn.argtypes = [C.c_double]
    L.rsk_ctx_create.argtypes = [C.c_int, C.POINTER(Params), C.c_void_p, C.POINTER(C.c_void_p)]
    L.rsk_ctx_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
    L.rsk_ctx_destroy.argtypes = [C.c_void_p]
    L.rsk_ctx_destroy.restype = None
    L.rsk_ctx_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.rsk_ctx_sync.argtypes = [C.c_void_p]
    L.rsk_chainset_upload.argtypes = [C.c_void_p, C.POINTER(ChainsHost), C.POINTER(C.c_void_p)]
    L.rsk_chainset_free.argtypes = [C.c_void_p]
    L.rsk_chainset_free.restype = None
    L.rsk_chainset_count.argtypes = [C.c_void_p]
    L.rsk_chainset_residues.argtypes = [C.c_void_p]
    L.rsk_search_cross.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SearchOpts), C.POINTER(C.c_void_p)]
    L.rsk_search_self.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(SearchOpts), C.POINTER(C.c_void_p)]
    L.rsk_search_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                   C.POINTER(SearchOpts), C.POINTER(C.c_void_p)]
    L.rsk_align_global.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.rsk_mu_gapless_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.rsk_chainset_selfrev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.rsk_search_cross_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SearchOpts)]
    for fn in (L.rsk_results_count, L.rsk_results_hits, L.rsk_results_paths, L.rsk_results_paths_bytes):
        fn.argtypes = [C.c_void_p]
    L.rsk_format_tsv.argtypes = [C.POINTER(HitView), C.c_int, C.c_char_p, C.c_char_p, C.c_size_t]
    L.rsk_path_to_cigar.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_char_p, C.c_size_t]
    L.rsk_kabsch.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int,
                             C.c_void_p, C.c_void_p, C.c_void_p]
    L.rsk_format_aln.argtypes = [C.POINTER(HitView), C.c_int, C.c_uint32, C.c_char_p, C.c_size_t]
    L.rsk_format_aln.restype = C.c_longlong
    L.rsk_format_fasta2.argtypes = [C.POINTER(HitView), C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    L.rsk_format_fasta2.restype = C.c_longlong
    L.rsk_results_free.argtypes = [C.c_void_p]
    L.rsk_results_free.restype = None
    L.rsk_prefilter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PrefilterOpts), C.POINTER(C.c_void_p)]
    for fn in (L.rsk_prefilter_count, L.rsk_prefilter_raw_count):
        fn.argtypes = [C.c_void_p]
        fn.restype = C.c_uint64
    for fn in (L.rsk_prefilter_targets, L.rsk_prefilter_queries, L.rsk_prefilter_scores):
        fn.argtypes = [C.c_void_p]
        fn.restype = C.c_void_p
    L.rsk_prefilter_bag.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
    L.rsk_prefilter_bag_device.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                           C.POINTER(C.c_void_p)]
    L.rsk_prefilter_select.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
    L.rsk_prefilter_free.argtypes = [C.c_void_p]
    L.rsk_prefilter_free.restype = None
    L.rsk_prefilter_to_tsv.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.rsk_prefilter_to_tsv.restype = C.c_longlong
    L.rsk_postfilter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SearchOpts), C.POINTER(C.c_void_p)]
    L.rsk_search_fast_db.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PrefilterOpts), C.POINTER(SearchOpts),
                                     C.POINTER(C.c_void_p)]
    L.rsk_chainset_from_coords.argtypes = [C.c_void_p, C.POINTER(CoordsHost), C.c_int, C.POINTER(C.c_void_p)]
    L.rsk_chainset_reversed.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.rsk_chainset_download_features.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.rsk_comm_unique_id.argtypes = [C.c_void_p]
    L.rsk_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.rsk_comm_create_all.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
    L.rsk_comm_destroy.argtypes = [C.c_void_p]
    L.rsk_comm_destroy.restype = None
    L.rsk_comm_rank.argtypes = [C.c_void_p]
    L.rsk_comm_nranks.argtypes = [C.c_void_p]
    L.rsk_comm_get_stats.argtypes = [C.c_void_p, C.POINTER(CommStats)]
    L.rsk_comm_reset_stats.argtypes = [C.c_void_p]
    L.rsk_partition_by_residues.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
    L.rsk_search_cross_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(SearchOpts),
                                           C.c_int, C.POINTER(C.c_void_p)]
    L.rsk_search_self_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SearchOpts), C.c_int, C.POINTER(C.c_void_p)]
    L.rsk_search_fast_db_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(PrefilterOpts),
                                             C.POINTER(SearchOpts), C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.rsk_results_digest.argtypes = [C.c_void_p]
    L.rsk_results_digest.restype = C.c_uint64
    _lib = L
    return L


def version():
    return load_library().rsk_version().decode()


def device_count():
    return load_library().rsk_device_count()


def _check(rc):
    if rc != 0:
        raise ReseekB200Error(f"libreseek_b200 error {rc}: {load_library().rsk_last_error().decode()}")


def params_preset(mode):
    p = Params()
    _check(load_library().rsk_params_preset(C.byref(p), int(mode)))
    return p


def format_tsv(hit, path, label_a, label_b, len_a, len_b, up=True, columns=None, seq_a=None, seq_b=None):
    """One TSV line for a hit record (numpy void of HIT_DTYPE), as DSSAligner::ToTsv would print it."""
    L = load_library()
    rec = np.array([hit], dtype=HIT_DTYPE)
    v = HitView(rec.ctypes.data, path.encode() if isinstance(path, str) else path, label_a.encode(), label_b.encode(),
                seq_a, seq_b, int(len_a), int(len_b))
    out = C.create_string_buffer(1 << 16)
    n = L.rsk_format_tsv(C.byref(v), int(bool(up)), columns.encode() if columns else None, out, 1 << 16)
    if n < 0:
        raise ReseekB200Error(f"rsk_format_tsv failed ({n})")
    return out.value.decode()


def _format_block(fn, name, hit, path, label_a, label_b, seq_a, seq_b, up, arg):
    rec = np.array([hit], dtype=HIT_DTYPE)
    v = HitView(rec.ctypes.data, path.encode() if isinstance(path, str) else path, label_a.encode(), label_b.encode(),
                seq_a, seq_b, len(seq_a), len(seq_b))
    cap = 1 << 14
    while True:
        out = C.create_string_buffer(cap)
        n = fn(C.byref(v), int(bool(up)), arg, out, cap)
        if n >= 0:
            return out.raw[:n].decode()
        if n >= -8:
            raise ReseekB200Error(f"{name} failed ({n})")
        cap = -n + 64


def format_aln(hit, path, label_a, label_b, seq_a, seq_b, up=True, rowlen=0):
    """The block DSSAligner::ToAln appends to the -aln file for this hit."""
    return _format_block(load_library().rsk_format_aln, "rsk_format_aln", hit, path, label_a, label_b, seq_a, seq_b, up, int(rowlen))


def format_fasta2(hit, path, label_a, label_b, seq_a, seq_b, up=True, unaligned=False):
    """The record DSSAligner::ToFasta2 appends to the -fasta2 file for this hit."""
    return _format_block(load_library().rsk_format_fasta2, "rsk_format_fasta2", hit, path, label_a, label_b, seq_a, seq_b, up,
                         int(bool(unaligned)))


def kabsch(xyz_a, xyz_b, lo_a, lo_b, path, up=True):
    """Superposition of the aligned residue pairs (DSSAligner::GetKabsch): returns (msd, t[3], u[3][3]) with y ~ u x + t."""
    xa = np.ascontiguousarray(xyz_a, np.float32)
    xb = np.ascontiguousarray(xyz_b, np.float32)
    t = np.zeros(3, np.float64)
    u = np.zeros(9, np.float64)
    msd = C.c_double()
    p = path.encode() if isinstance(path, str) else path
    _check(load_library().rsk_kabsch(xa.ctypes.data, xa.shape[1], xb.ctypes.data, xb.shape[1], int(lo_a), int(lo_b), p, len(p),
                                     int(bool(up)), t.ctypes.data, u.ctypes.data, C.addressof(msd)))
    return msd.value, t, u.reshape(3, 3)


def path_to_cigar(path, up=True):
    L = load_library()
    out = C.create_string_buffer(4 * len(path) + 16)
    n = L.rsk_path_to_cigar(path.encode(), len(path), int(bool(up)), out, len(out))
    if n < 0:
        raise ReseekB200Error(f"rsk_path_to_cigar failed ({n})")
    return out.value.decode()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Results:
    """Hits of one search call.  `hits` is a zero-copy numpy view (HIT_DTYPE) of the library's rsk_hit array and
    `paths` a view of its path pool; both stay valid while this object is alive."""

    def __init__(self, handle):
        L = load_library()
        self._handle = handle
        n = L.rsk_results_count(handle)
        if n:
            buf = (C.c_char * (n * HIT_DTYPE.itemsize)).from_address(L.rsk_results_hits(handle))
            self.hits = np.frombuffer(buf, dtype=HIT_DTYPE, count=n)
        else:
            self.hits = np.zeros(0, HIT_DTYPE)
        nb = L.rsk_results_paths_bytes(handle)
        if nb:
            buf = (C.c_char * nb).from_address(L.rsk_results_paths(handle))
            self.paths = np.frombuffer(buf, dtype=np.uint8, count=nb)
        else:
            self.paths = np.zeros(0, np.uint8)

    def path(self, k):
        h = self.hits[k]
        off, n = int(h["path_off"]), int(h["path_len"])
        return self.paths[off:off + n].tobytes().decode()

    def close(self):
        if self._handle:
            self.hits = np.zeros(0, HIT_DTYPE)
            self.paths = np.zeros(0, np.uint8)
            load_library().rsk_results_free(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return len(self.hits)

    def digest(self):
        """Order-independent 64-bit digest of records + paths (rsk_results_digest)."""
        return int(load_library().rsk_results_digest(self._handle))


def partition_by_residues(lens, nranks):
    """Contiguous blocks with (nearly) equal residue totals (rsk_partition_by_residues): list of (lo, hi) per rank."""
    lens = np.ascontiguousarray(lens, np.uint32)
    b = np.zeros(nranks + 1, np.uint32)
    _check(load_library().rsk_partition_by_residues(_ptr(lens), len(lens), int(nranks), _ptr(b)))
    return [(int(b[r]), int(b[r + 1])) for r in range(nranks)]


def comm_unique_id():
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(load_library().rsk_comm_unique_id(buf))
    return buf.raw


class Comm:
    """One rank's communicator (NCCL over NVLink) for the DB-sharded searches; nranks == 1 needs no id."""

    def __init__(self, ctx, nranks=1, rank=0, unique_id=None, handle=None):
        self.ctx = ctx
        if handle is not None:
            self.handle = handle
            return
        self.handle = C.c_void_p()
        idbuf = C.create_string_buffer(unique_id, COMM_ID_BYTES) if unique_id is not None else None
        _check(load_library().rsk_comm_create(ctx.handle, int(nranks), int(rank), idbuf, C.byref(self.handle)))

    @classmethod
    def from_torch_dist(cls, ctx, dist):
        """Rank/size from an initialised torch.distributed group; the NCCL id is made on rank 0 and broadcast through it."""
        import torch
        world, rank = dist.get_world_size(), dist.get_rank()
        if world == 1:
            return cls(ctx, 1, 0)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.zeros(COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return cls(ctx, world, rank, bytes(t.cpu().numpy().tobytes()))

    @classmethod
    def create_all(cls, ctxs):
        """One process driving several GPUs: communicators for contexts on different devices (ncclCommInitAll)."""
        n = len(ctxs)
        arr = (C.c_void_p * n)(*[c.handle for c in ctxs])
        out = (C.c_void_p * n)()
        _check(load_library().rsk_comm_create_all(arr, n, out))
        return [cls(ctxs[k], handle=C.c_void_p(out[k])) for k in range(n)]

    @property
    def rank(self):
        return load_library().rsk_comm_rank(self.handle)

    @property
    def nranks(self):
        return load_library().rsk_comm_nranks(self.handle)

    def stats(self):
        s = CommStats()
        _check(load_library().rsk_comm_get_stats(self.handle, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        _check(load_library().rsk_comm_reset_stats(self.handle))

    def close(self):
        if self.handle:
            load_library().rsk_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PrefilterResult:
    """Candidate (target, query, score) lists of rsk_prefilter, in the order of the reference's candidate TSV."""

    def __init__(self, handle):
        L = load_library()
        self._handle = handle
        n = L.rsk_prefilter_count(handle)
        self.raw_count = int(L.rsk_prefilter_raw_count(handle))

        def view(fn, dt):
            if n == 0:
                return np.zeros(0, dt)
            buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(fn(handle))
            return np.frombuffer(buf, dtype=dt, count=n).copy()
        self.targets = view(L.rsk_prefilter_targets, np.uint32)
        self.queries = view(L.rsk_prefilter_queries, np.uint32)
        self.scores = view(L.rsk_prefilter_scores, np.uint16)

    def tsv(self):
        L = load_library()
        need = -L.rsk_prefilter_to_tsv(self._handle, None, 0)
        buf = C.create_string_buffer(int(need))
        n = L.rsk_prefilter_to_tsv(self._handle, buf, int(need))
        if n < 0:
            raise ReseekB200Error("rsk_prefilter_to_tsv failed")
        return buf.value.decode()

    def select(self, t_lo, t_hi):
        """Candidates with t_lo <= target < t_hi, re-based to t_lo (one rank's share of a merged list)."""
        r = C.c_void_p()
        _check(load_library().rsk_prefilter_select(self._handle, int(t_lo), int(t_hi), C.byref(r)))
        return PrefilterResult(r)

    def as_dict(self):
        out = {}
        for t, q in zip(self.targets.tolist(), self.queries.tolist()):
            out.setdefault(t, []).append(q)
        return out

    def close(self):
        if self._handle:
            load_library().rsk_prefilter_free(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return len(self.targets)


def prefilter_bag(nq, targets, queries, scores, rsb_size=0):
    """RankedScoresBag over (target, query, score) triples in stream order (rsk_prefilter_bag; host only)."""
    t = np.ascontiguousarray(targets, np.uint32)
    q = np.ascontiguousarray(queries, np.uint32)
    s = np.ascontiguousarray(scores, np.uint16)
    assert len(t) == len(q) == len(s)
    r = C.c_void_p()
    _check(load_library().rsk_prefilter_bag(int(nq), len(t), _ptr(t), _ptr(q), _ptr(s), int(rsb_size), C.byref(r)))
    return PrefilterResult(r)


class ChainSet:
    """Device-resident chains (DBSearcher's per-chain vectors, dbsearcher.h:26-33)."""

    def __init__(self, ctx, lens, prof, mu, xyz, selfrev):
        self.ctx = ctx
        self.lens = np.ascontiguousarray(lens, np.uint32)
        self.n = len(self.lens)
        self.total = int(self.lens.sum(dtype=np.uint64))
        self.prof = np.ascontiguousarray(prof, np.uint8)
        assert self.prof.shape == (NFEAT, self.total), (self.prof.shape, self.total)
        self.mu = None if mu is None else np.ascontiguousarray(mu, np.uint8)
        self.xyz = np.ascontiguousarray(xyz, np.float32)
        assert self.xyz.shape == (3, self.total)
        self.selfrev = None if selfrev is None else np.ascontiguousarray(selfrev, np.float32)
        h = ChainsHost(self.n, self.total, _ptr(self.lens), _ptr(self.prof), _ptr(self.mu), _ptr(self.xyz),
                       _ptr(self.selfrev))
        self.handle = C.c_void_p()
        _check(load_library().rsk_chainset_upload(ctx.handle, C.byref(h), C.byref(self.handle)))

    @classmethod
    def _from_handle(cls, ctx, handle, lens, has_mu):
        self = cls.__new__(cls)
        self.ctx = ctx
        self.handle = handle
        self.lens = np.ascontiguousarray(lens, np.uint32)
        self.n = len(self.lens)
        self.total = int(self.lens.sum(dtype=np.uint64))
        self.prof = self.xyz = self.selfrev = None
        self.mu = True if has_mu else None
        return self

    @classmethod
    def from_coords(cls, ctx, lens, aa, xyz, with_mu=True):
        """DSS on the device (rsk_chainset_from_coords): aa = amino-acid characters [total] (bytes / uint8), xyz [3][total]."""
        lens = np.ascontiguousarray(lens, np.uint32)
        aa = np.ascontiguousarray(np.frombuffer(aa, np.uint8) if isinstance(aa, (bytes, bytearray)) else aa, np.uint8)
        xyz = np.ascontiguousarray(xyz, np.float32)
        total = int(lens.sum(dtype=np.uint64))
        assert len(aa) == total and xyz.shape == (3, total)
        h = CoordsHost(len(lens), total, _ptr(lens), _ptr(aa), _ptr(xyz))
        handle = C.c_void_p()
        _check(load_library().rsk_chainset_from_coords(ctx.handle, C.byref(h), int(bool(with_mu)), C.byref(handle)))
        return cls._from_handle(ctx, handle, lens, with_mu)

    def reversed(self):
        """PDBChain::GetReverse + DSS of every chain on the device, forward Mu letters (rsk_chainset_reversed)."""
        handle = C.c_void_p()
        _check(load_library().rsk_chainset_reversed(self.ctx.handle, self.handle, C.byref(handle)))
        return ChainSet._from_handle(self.ctx, handle, self.lens, self.mu is not None)

    def download_features(self, want_mu=True):
        """(prof [8][total], mu [total] or None, selfrev [n]) read back from the device."""
        prof = np.zeros((NFEAT, self.total), np.uint8)
        mu = np.zeros(self.total, np.uint8) if (want_mu and self.mu is not None) else None
        sr = np.zeros(self.n, np.float32)
        _check(load_library().rsk_chainset_download_features(self.ctx.handle, self.handle, _ptr(prof), _ptr(mu), _ptr(sr)))
        return prof, mu, sr

    @property
    def h2d_bytes(self):
        if self.prof is None:
            return 13 * self.total + 4 * self.n
        return self.lens.nbytes + self.prof.nbytes + (0 if self.mu is None else self.mu.nbytes) + self.xyz.nbytes + \
            (0 if self.selfrev is None else self.selfrev.nbytes)

    def free(self):
        if self.handle:
            load_library().rsk_chainset_free(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU + stream + scratch: the aligner pool of a DBSearcher (dbsearcher.cpp:73)."""

    def __init__(self, device=0, mode=MODE_VERYSENSITIVE, params=None, stream=None):
        L = load_library()
        self.params = params if params is not None else params_preset(mode)
        self.handle = C.c_void_p()
        _check(L.rsk_ctx_create(int(device), C.byref(self.params), C.c_void_p(stream) if stream else None,
                                C.byref(self.handle)))

    def set_params(self, params):
        self.params = params
        _check(load_library().rsk_ctx_set_params(self.handle, C.byref(params)))

    def upload(self, lens, prof, mu, xyz, selfrev=None):
        return ChainSet(self, lens, prof, mu, xyz, selfrev)

    def upload_chains(self, chains):
        """chains: objects with .prof [8][L], .mu, .xyz [3][L], .selfrev (duck-typed)."""
        lens = np.array([c.L for c in chains], np.uint32)
        prof = np.concatenate([c.prof for c in chains], axis=1)
        mu = None if any(c.mu is None for c in chains) else np.concatenate([c.mu for c in chains])
        xyz = np.concatenate([c.xyz for c in chains], axis=1)
        selfrev = np.array([c.selfrev for c in chains], np.float32)
        return self.upload(lens, prof, mu, xyz, selfrev)

    @staticmethod
    def _opts(keep, want_paths, skip_evalue):
        return SearchOpts(int(keep), int(bool(want_paths)), int(bool(skip_evalue)), 0)

    def search_cross(self, A, B, keep=KEEP_ALL, want_paths=True, skip_evalue=False):
        o = self._opts(keep, want_paths, skip_evalue)
        r = C.c_void_p()
        _check(load_library().rsk_search_cross(self.handle, A.handle, B.handle, C.byref(o), C.byref(r)))
        return Results(r)

    def search_cross_device(self, A, B, skip_evalue=False):
        o = self._opts(KEEP_ALL, False, skip_evalue)
        _check(load_library().rsk_search_cross_device(self.handle, A.handle, B.handle, C.byref(o)))

    def search_self(self, S, keep=KEEP_ALL, want_paths=True, skip_evalue=False):
        o = self._opts(keep, want_paths, skip_evalue)
        r = C.c_void_p()
        _check(load_library().rsk_search_self(self.handle, S.handle, C.byref(o), C.byref(r)))
        return Results(r)

    def search_pairs(self, A, B, ia, ib, keep=KEEP_ALL, want_paths=True, skip_evalue=False):
        ia = np.ascontiguousarray(ia, np.uint32)
        ib = np.ascontiguousarray(ib, np.uint32)
        assert len(ia) == len(ib)
        o = self._opts(keep, want_paths, skip_evalue)
        r = C.c_void_p()
        _check(load_library().rsk_search_pairs(self.handle, A.handle, B.handle, len(ia), _ptr(ia), _ptr(ib),
                                               C.byref(o), C.byref(r)))
        return Results(r)

    def align_global(self, A, B, ia, ib):
        """-global: DSSAligner::AlignQueryTarget_Global (global.cpp:7-33) for explicit pairs; one record per pair, in order,
        score = m_GlobalScore, whole-chain path."""
        ia = np.ascontiguousarray(ia, np.uint32)
        ib = np.ascontiguousarray(ib, np.uint32)
        assert len(ia) == len(ib)
        r = C.c_void_p()
        _check(load_library().rsk_align_global(self.handle, A.handle, B.handle, len(ia), _ptr(ia), _ptr(ib), C.byref(r)))
        return Results(r)

    def prefilter(self, Q, T, index_mode=0, rsb_size=0, kl_swap=True, raw_only=False):
        """MuPreFilter (muprefilter.cpp:64-133) on the GPU; Q = -search chains, T = -db chains.
        raw_only: every (target, query, score) triple with a two-hit diagonal, before the per-query bag."""
        o = PrefilterOpts(int(index_mode), int(rsb_size), int(not kl_swap), int(bool(raw_only)))
        r = C.c_void_p()
        _check(load_library().rsk_prefilter(self.handle, Q.handle, T.handle, C.byref(o), C.byref(r)))
        return PrefilterResult(r)

    def postfilter(self, Q, T, cands, keep=KEEP_HITS, want_paths=True):
        o = self._opts(keep, want_paths, False)
        r = C.c_void_p()
        _check(load_library().rsk_postfilter(self.handle, Q.handle, T.handle, cands._handle, C.byref(o), C.byref(r)))
        return Results(r)

    def search_fast_db(self, Q, T, index_mode=0, rsb_size=0, kl_swap=True, keep=KEEP_HITS, want_paths=True):
        """`reseek -search Q -db T -fast` (search.cpp:76-111): prefilter + post-filter; hit.a = query, hit.b = target."""
        po = PrefilterOpts(int(index_mode), int(rsb_size), int(not kl_swap), 0)
        o = self._opts(keep, want_paths, False)
        r = C.c_void_p()
        _check(load_library().rsk_search_fast_db(self.handle, Q.handle, T.handle, C.byref(po), C.byref(o), C.byref(r)))
        return Results(r)

    def search_cross_sharded(self, comm, A_local, B, a_base, keep=KEEP_HITS, want_paths=True, root=0, skip_evalue=False):
        """DBSearcher::RunQuery over this rank's block of the -db chains (rsk_search_cross_sharded): hits compacted on the
        device and gathered on `root` over NVLink; returns Results on the root, None elsewhere.  comm=None: one rank."""
        o = self._opts(keep, want_paths, skip_evalue)
        r = C.c_void_p()
        _check(load_library().rsk_search_cross_sharded(self.handle, comm.handle if comm else None,
                                                        A_local.handle if A_local is not None else None, B.handle, int(a_base),
                                                        C.byref(o), int(root), C.byref(r)))
        return Results(r) if r else None

    def search_self_sharded(self, comm, S, keep=KEEP_HITS, want_paths=True, root=0):
        """DBSearcher::RunSelf with the rows of the pair triangle interleaved over the ranks (rsk_search_self_sharded)."""
        o = self._opts(keep, want_paths, False)
        r = C.c_void_p()
        _check(load_library().rsk_search_self_sharded(self.handle, comm.handle if comm else None, S.handle, C.byref(o), int(root), C.byref(r)))
        return Results(r) if r else None

    def search_fast_db_sharded(self, comm, Q, T_local, t_base, index_mode=0, rsb_size=0, kl_swap=True, keep=KEEP_HITS,
                               want_paths=True, root=0, want_cands=True):
        """`-search Q -db DB -fast` on this rank's block of the DB (rsk_search_fast_db_sharded).  Returns
        (Results on the root / None, merged candidate list or None)."""
        po = PrefilterOpts(int(index_mode), int(rsb_size), int(not kl_swap), 0)
        o = self._opts(keep, want_paths, False)
        r, c = C.c_void_p(), C.c_void_p()
        _check(load_library().rsk_search_fast_db_sharded(self.handle, comm.handle if comm else None, Q.handle,
                                                          T_local.handle if T_local is not None else None, int(t_base),
                                                          C.byref(po), C.byref(o), int(root), C.byref(r),
                                                          C.byref(c) if want_cands else None))
        return (Results(r) if r else None), (PrefilterResult(c) if c else None)

    def prefilter_bag_device(self, nq, targets, queries, scores, rsb_size=0):
        """RankedScoresBag on the device over (target, query, score) triples in stream order (rsk_prefilter_bag_device)."""
        t = np.ascontiguousarray(targets, np.uint32)
        q = np.ascontiguousarray(queries, np.uint32)
        s = np.ascontiguousarray(scores, np.uint16)
        assert len(t) == len(q) == len(s)
        r = C.c_void_p()
        _check(load_library().rsk_prefilter_bag_device(self.handle, int(nq), len(t), _ptr(t), _ptr(q), _ptr(s), int(rsb_size), C.byref(r)))
        return PrefilterResult(r)

    def selfrev(self, S, Srev):
        """GetSelfRevScore for every chain of S (see rsk_chainset_selfrev); returns the float32 scores."""
        out = np.zeros(S.n, np.float32)
        _check(load_library().rsk_chainset_selfrev(self.handle, S.handle, Srev.handle, _ptr(out)))
        return out

    def mu_gapless_scores(self, A, B, ia, ib):
        """(SWFastGaplessProfb float scores, SWFastPinopGapless int scores) for the listed pairs."""
        ia = np.ascontiguousarray(ia, np.uint32)
        ib = np.ascontiguousarray(ib, np.uint32)
        f = np.zeros(len(ia), np.float32)
        i = np.zeros(len(ia), np.int32)
        _check(load_library().rsk_mu_gapless_scores(self.handle, A.handle, B.handle, len(ia), _ptr(ia), _ptr(ib), _ptr(f), _ptr(i)))
        return f, i

    def stats(self):
        s = Stats()
        _check(load_library().rsk_ctx_stats(self.handle, C.byref(s)))
        return s.as_dict()

    def sync(self):
        _check(load_library().rsk_ctx_sync(self.handle))

    def close(self):
        if self.handle:
            load_library().rsk_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
