"""ctypes bindings for the oracle libraries.  TEST INFRASTRUCTURE ONLY.

  Port  - oracle/_port/liboracle_port.so : the plain-C restatement (oracle/reseek_oracle.c)
  Ref   - oracle/_ref/libreseek_ref.so   : the unmodified reference behind oracle/ref_driver.cpp

Only tests/, tools/make_golden.py, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package reseek_b200 never does.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PORT_SO = HERE / "_port" / "liboracle_port.so"
REF_SO = HERE / "_ref" / "libreseek_ref.so"
REF_FAST_SO = HERE / "_ref" / "libreseek_ref_fast.so"  # -O3 build of the same sources: timed only, never used for parity
REF_BIN = HERE / "_ref" / "reseek_ref"

NFEAT = 8
FLT_MAX = float(np.finfo(np.float32).max)
U32_MAX = 0xFFFFFFFF

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build_port():
    subprocess.run(["make", "-C", str(HERE), "port"], check=True, capture_output=True)


def build_ref():
    subprocess.run(["make", "-C", str(HERE), "-j8", "ref"], check=True, capture_output=True)


class OrcParams(C.Structure):
    _fields_ = [("gap_open", C.c_float), ("gap_ext", C.c_float), ("min_fwd_score", C.c_float),
                ("omega", C.c_float), ("omega_fwd", C.c_float), ("mu_gap_open", C.c_int),
                ("mu_gap_ext", C.c_int), ("mkfl", C.c_uint32), ("mkf_x1", C.c_int), ("mkf_x2", C.c_int),
                ("mkf_min_hsp_score", C.c_int), ("mkf_min_mega_hsp_score", C.c_float),
                ("weights", C.c_float * NFEAT), ("tables", C.c_float * 2192)]


class OrcChain(C.Structure):
    _fields_ = [("L", C.c_uint32), ("prof", C.c_void_p), ("mu", C.c_void_p), ("x", C.c_void_p),
                ("y", C.c_void_p), ("z", C.c_void_p), ("selfrev", C.c_float)]


class OrcResult(C.Structure):
    _fields_ = [("score", C.c_float), ("lo_a", C.c_uint32), ("lo_b", C.c_uint32), ("hi_a", C.c_uint32),
                ("hi_b", C.c_uint32), ("ids", C.c_uint32), ("gaps", C.c_uint32), ("lddt", C.c_float),
                ("ts", C.c_float), ("pvalue", C.c_float), ("evalue", C.c_float), ("qual", C.c_float),
                ("mu_score", C.c_float), ("mu_fwd", C.c_int32), ("mu_rev", C.c_int32),
                ("filtered", C.c_int32), ("path_len", C.c_uint32)]


class RefResult(C.Structure):
    _fields_ = [("score", C.c_float), ("lo_a", C.c_uint32), ("lo_b", C.c_uint32), ("hi_a", C.c_uint32),
                ("hi_b", C.c_uint32), ("ids", C.c_uint32), ("gaps", C.c_uint32), ("lddt", C.c_float),
                ("ts", C.c_float), ("pvalue", C.c_float), ("evalue", C.c_float), ("qual", C.c_float),
                ("mu_score", C.c_float), ("mkf", C.c_int32), ("best_hsp_score", C.c_int32),
                ("best_chain_score", C.c_int32), ("xdrop_score", C.c_float), ("path_len", C.c_uint32)]


class Chain:
    """One chain as the aligner sees it: 8 feature planes, Mu letters, coordinates, self-reverse score."""

    def __init__(self, prof, mu=None, xyz=None, selfrev=FLT_MAX, label="", seq=None, kmers=None):
        self.prof = np.ascontiguousarray(prof, dtype=np.uint8)  # [8][L]
        self.L = int(self.prof.shape[1])
        self.mu = None if mu is None else np.ascontiguousarray(mu, dtype=np.uint8)
        if xyz is None:
            xyz = np.zeros((3, self.L), np.float32)
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float32)  # [3][L]
        self.selfrev = float(selfrev)
        self.label = label
        self.seq = seq
        self.kmers = None if kmers is None else np.ascontiguousarray(kmers, dtype=np.uint32)

    def as_orc(self):
        c = OrcChain()
        c.L = self.L
        c.prof = self.prof.ctypes.data
        c.mu = None if self.mu is None else self.mu.ctypes.data
        c.x = self.xyz[0].ctypes.data
        c.y = self.xyz[1].ctypes.data
        c.z = self.xyz[2].ctypes.data
        c.selfrev = self.selfrev
        return c


class Port:
    """The plain-C oracle (oracle/reseek_oracle.c)."""

    def __init__(self, mode=3):
        if not PORT_SO.exists():
            build_port()
        L = self.lib = C.CDLL(str(PORT_SO))
        L.orc_sw_align.restype = C.c_float
        L.orc_swfast_matrix.restype = C.c_float
        L.orc_cell_score.restype = C.c_float
        L.orc_mu_filter_score.restype = C.c_float
        L.orc_lddt.restype = C.c_float
        L.orc_mu_gapless_profb.restype = C.c_float
        L.orc_pvalue.restype = C.c_double
        L.orc_evalue.restype = C.c_double
        L.orc_qual.restype = C.c_double
        L.orc_pvalue.argtypes = L.orc_evalue.argtypes = L.orc_qual.argtypes = [C.c_double]
        L.orc_bgfreq.restype = C.POINTER(C.c_float)
        L.orc_mu_i8.restype = C.POINTER(C.c_int8)
        L.orc_mu_kmer_i8.restype = C.POINTER(C.c_int8)
        L.orc_mu_f32.restype = C.POINTER(C.c_float)
        self.params = OrcParams()
        self.set_mode(mode)

    def set_mode(self, mode):
        assert self.lib.orc_params_preset(C.byref(self.params), int(mode)) == 0
        self.mode = mode

    def tables(self):
        return np.array(self.params.tables[:], dtype=np.float32)

    def feat_alpha(self):
        return [self.lib.orc_feat_alpha(f) for f in range(NFEAT)]

    def bgfreq(self, f):
        n = self.lib.orc_feat_alpha(f)
        return np.array(self.lib.orc_bgfreq(f)[:n], dtype=np.float64)

    def mu_i8(self):
        return np.array(self.lib.orc_mu_i8()[:1296], dtype=np.int8).reshape(36, 36)

    def mu_kmer_i8(self):
        return np.array(self.lib.orc_mu_kmer_i8()[:1296], dtype=np.int8).reshape(36, 36)

    def mu_f32(self):
        return np.array(self.lib.orc_mu_f32()[:1296], dtype=np.float32).reshape(36, 36)

    def sw_align(self, profA, profB):
        profA = np.ascontiguousarray(profA, np.uint8)
        profB = np.ascontiguousarray(profB, np.uint8)
        LA, LB = profA.shape[1], profB.shape[1]
        lo_a, lo_b, plen = C.c_uint32(), C.c_uint32(), C.c_uint32()
        path = C.create_string_buffer(LA + LB + 2)
        s = self.lib.orc_sw_align(C.byref(self.params), profA.ctypes.data_as(C.c_void_p), LA,
                                  profB.ctypes.data_as(C.c_void_p), LB, C.byref(lo_a), C.byref(lo_b), path,
                                  C.byref(plen))
        return float(np.float32(s)), lo_a.value, lo_b.value, path.value.decode()

    def swfast_matrix(self, S, open_, ext):
        S = np.ascontiguousarray(S, np.float32)
        LA, LB = S.shape
        lo_a, lo_b, plen = C.c_uint32(), C.c_uint32(), C.c_uint32()
        path = C.create_string_buffer(LA + LB + 2)
        s = self.lib.orc_swfast_matrix(S.ctypes.data_as(C.c_void_p), LA, LB, C.c_float(open_), C.c_float(ext),
                                       C.byref(lo_a), C.byref(lo_b), path, C.byref(plen))
        return float(np.float32(s)), lo_a.value, lo_b.value, path.value.decode()

    def score_matrix(self, profA, profB):
        profA = np.ascontiguousarray(profA, np.uint8)
        profB = np.ascontiguousarray(profB, np.uint8)
        LA, LB = profA.shape[1], profB.shape[1]
        S = np.empty((LA, LB), np.float32)
        for i in range(LA):
            for j in range(LB):
                S[i, j] = self.lib.orc_cell_score(C.byref(self.params), profA.ctypes.data_as(C.c_void_p), LA, i,
                                                  profB.ctypes.data_as(C.c_void_p), LB, j)
        return S

    def mu_sw(self, a, b, open_=2, ext=1):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        sat = C.c_int()
        s = self.lib.orc_mu_sw_score(a.ctypes.data_as(C.c_void_p), len(a), b.ctypes.data_as(C.c_void_p), len(b),
                                     open_, ext, C.byref(sat))
        return s, sat.value

    def mu_filter_score(self, a, b):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        fwd, rev = C.c_int(), C.c_int()
        s = self.lib.orc_mu_filter_score(C.byref(self.params), a.ctypes.data_as(C.c_void_p), len(a),
                                         b.ctypes.data_as(C.c_void_p), len(b), C.byref(fwd), C.byref(rev))
        return float(s), fwd.value, rev.value

    def mu_gapless(self, a, b):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        pa, pb = a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)
        return (float(np.float32(self.lib.orc_mu_gapless_profb(pa, len(a), pb, len(b)))),
                int(self.lib.orc_mu_gapless_int(pa, len(a), pb, len(b))))

    def prefilter(self, mu_queries, mu_targets, query_neighborhood=True, rsb_size=1500, swap_kl=True):
        """MuPreFilter (muprefilter.cpp:64-133): returns {target index: [query indices]} as written to the
        prefilter TSV (rankedscoresbag.cpp:185-232), plus the raw per-(target, query) diagonal scores."""
        L = self.lib
        L.orc_rsb_new.restype = C.c_void_p
        L.orc_rsb_targets.restype = C.POINTER(C.c_uint32)
        L.orc_rsb_scores.restype = C.POINTER(C.c_uint16)
        qs = []
        for m in mu_queries:
            m = np.array(m, np.uint8)
            if swap_kl:  # the query side goes through g_CharToLetterMu, which exchanges letters 10 and 11
                k, l = m == 10, m == 11
                m[k], m[l] = 11, 10
            qs.append(np.ascontiguousarray(m))
        nQ = len(qs)
        qptr = (C.c_void_p * nQ)(*[q.ctypes.data for q in qs])
        LQ = np.array([len(q) for q in qs], np.uint32)
        rsb = C.c_void_p(L.orc_rsb_new(nQ, int(rsb_size)))
        best = np.zeros(nQ, np.uint16)
        raw = {}
        for t, mt in enumerate(mu_targets):
            mt = np.ascontiguousarray(mt, np.uint8)
            if len(mt) == 0:
                continue
            L.orc_prefilter_target(qptr, LQ.ctypes.data_as(C.c_void_p), nQ, mt.ctypes.data_as(C.c_void_p), len(mt),
                                   int(query_neighborhood), best.ctypes.data_as(C.c_void_p))
            for q in np.nonzero(best)[0]:
                raw[(t, int(q))] = int(best[q])
                L.orc_rsb_add(rsb, int(q), t, int(best[q]))
        L.orc_rsb_finish(rsb)
        out = {}
        for q in range(nQ):
            n = L.orc_rsb_count(rsb, q)
            tg = L.orc_rsb_targets(rsb, q)
            for k in range(n):
                out.setdefault(int(tg[k]), []).append(q)
        L.orc_rsb_free(rsb)
        return out, raw

    def lddt(self, A, B, posA, posB):
        posA = np.ascontiguousarray(posA, np.uint32)
        posB = np.ascontiguousarray(posB, np.uint32)
        ca, cb = A.as_orc(), B.as_orc()
        return float(np.float32(self.lib.orc_lddt(C.byref(ca), C.byref(cb), posA.ctypes.data_as(C.c_void_p),
                                                  posB.ctypes.data_as(C.c_void_p), len(posA))))

    def statsig(self, ts):
        return self.lib.orc_pvalue(ts), self.lib.orc_evalue(ts), self.lib.orc_qual(ts)

    def align_pair(self, A, B):
        ca, cb = A.as_orc(), B.as_orc()
        r = OrcResult()
        path = C.create_string_buffer(A.L + B.L + 2)
        self.lib.orc_align_pair(C.byref(self.params), C.byref(ca), C.byref(cb), C.byref(r), path)
        return r, path.value.decode()

    def align_pair_global(self, A, B):
        """DSSAligner::AlignQueryTarget_Global: r.score = m_GlobalScore, lo = 0, whole-chain path."""
        ca, cb = A.as_orc(), B.as_orc()
        r = OrcResult()
        path = C.create_string_buffer(A.L + B.L + 2)
        self.lib.orc_align_pair_global(C.byref(self.params), C.byref(ca), C.byref(cb), C.byref(r), path)
        return r, path.value.decode()

    def align_pairs(self, chainsA, chainsB, ia, ib):
        """Batch (scalar, one thread) - used for the cpu_baseline timing.  Returns the OrcResult array."""
        arrA = (OrcChain * len(chainsA))(*[c.as_orc() for c in chainsA])
        arrB = (OrcChain * len(chainsB))(*[c.as_orc() for c in chainsB])
        ia = np.ascontiguousarray(ia, np.uint32)
        ib = np.ascontiguousarray(ib, np.uint32)
        out = (OrcResult * len(ia))()
        self.lib.orc_align_pairs(C.byref(self.params), arrA, arrB, ia.ctypes.data_as(C.c_void_p),
                                 ib.ctypes.data_as(C.c_void_p), C.c_size_t(len(ia)), out)
        return out


class Ref:
    """The unmodified reference (oracle/_ref/libreseek_ref.so) behind oracle/ref_driver.cpp."""

    _inited_mode = None

    def __init__(self, mode=3, fast=False):
        so = REF_FAST_SO if fast else REF_SO
        if not so.exists():
            raise FileNotFoundError(f"{so} missing - run `make -C oracle ref{'_fast' if fast else ''}` in the build container")
        L = self.lib = C.CDLL(str(so))
        L.ref_selfrev.restype = C.c_float
        L.ref_mu_score.restype = C.c_float
        L.ref_swfast.restype = C.c_float
        L.ref_gapless_profb.restype = C.c_float
        L.ref_lddt.restype = C.c_double
        L.ref_init(int(mode))
        self.mode = mode

    @staticmethod
    def available(fast=False):
        return (REF_FAST_SO if fast else REF_SO).exists()

    def get_params(self):
        sc = np.zeros(12, np.float32)
        tb = np.zeros(2192, np.float32)
        self.lib.ref_get_params(sc.ctypes.data_as(C.c_void_p), tb.ctypes.data_as(C.c_void_p))
        return sc, tb

    def mu_matrices(self):
        f = np.zeros(1296, np.float32)
        a = np.zeros(1296, np.int8)
        b = np.zeros(1296, np.int8)
        self.lib.ref_get_mu_matrices(f.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                                     b.ctypes.data_as(C.c_void_p))
        return f.reshape(36, 36), a.reshape(36, 36), b.reshape(36, 36)

    def bca_open(self, fn):
        return self.lib.ref_bca_open(str(fn).encode())

    def bca_chain(self, idx):
        L = self.lib.ref_bca_len(idx)
        label = C.create_string_buffer(512)
        seq = C.create_string_buffer(L + 1)
        xyz = np.zeros((3, L), np.float32)
        self.lib.ref_bca_chain(idx, label, 512, seq, xyz[0].ctypes.data_as(C.c_void_p),
                               xyz[1].ctypes.data_as(C.c_void_p), xyz[2].ctypes.data_as(C.c_void_p))
        return label.value.decode(), seq.raw[:L], xyz

    def dss(self, seq, xyz):
        L = xyz.shape[1]
        prof = np.zeros((8, L), np.uint8)
        mu = np.zeros(L, np.uint8)
        km = np.zeros(max(L, 1), np.uint32)
        nk = C.c_uint32()
        self.lib.ref_dss(L, seq, xyz[0].ctypes.data_as(C.c_void_p), xyz[1].ctypes.data_as(C.c_void_p),
                         xyz[2].ctypes.data_as(C.c_void_p), prof.ctypes.data_as(C.c_void_p),
                         mu.ctypes.data_as(C.c_void_p), km.ctypes.data_as(C.c_void_p), C.byref(nk))
        return prof, mu, km[:nk.value].copy()

    def selfrev(self, seq, xyz, loader=True, with_mu=True):
        L = xyz.shape[1]
        return float(np.float32(self.lib.ref_selfrev(L, seq, xyz[0].ctypes.data_as(C.c_void_p),
                                                     xyz[1].ctypes.data_as(C.c_void_p),
                                                     xyz[2].ctypes.data_as(C.c_void_p), int(loader), int(with_mu))))

    def rev_profile(self, seq, xyz):
        L = xyz.shape[1]
        prof = np.zeros((8, L), np.uint8)
        self.lib.ref_rev_profile(L, seq, xyz[0].ctypes.data_as(C.c_void_p), xyz[1].ctypes.data_as(C.c_void_p),
                                 xyz[2].ctypes.data_as(C.c_void_p), prof.ctypes.data_as(C.c_void_p))
        return prof

    def load_chain(self, idx, loader_selfrev=True):
        """Chain idx of the open .bca with features and self-reverse score from the reference itself."""
        label, seq, xyz = self.bca_chain(idx)
        prof, mu, km = self.dss(seq, xyz)
        sr = self.selfrev(seq, xyz, loader=loader_selfrev)
        return Chain(prof, mu, xyz, sr, label=label, seq=seq, kmers=km)

    def align_pair(self, A, B, noaccel=False, use_mu=True, use_kmers=True):
        r = RefResult()
        path = C.create_string_buffer(A.L + B.L + 2)

        def p(a):
            return None if a is None else a.ctypes.data_as(C.c_void_p)

        muA = A.mu if use_mu else None
        muB = B.mu if use_mu else None
        kA = A.kmers if (use_kmers and use_mu) else None
        kB = B.kmers if (use_kmers and use_mu) else None
        self.lib.ref_align_pair(
            A.L, p(A.prof), p(muA), p(kA), 0 if kA is None else len(kA), p(A.xyz[0]), p(A.xyz[1]), p(A.xyz[2]),
            C.c_float(A.selfrev),
            B.L, p(B.prof), p(muB), p(kB), 0 if kB is None else len(kB), p(B.xyz[0]), p(B.xyz[1]), p(B.xyz[2]),
            C.c_float(B.selfrev), int(noaccel), C.byref(r), path, A.L + B.L + 2)
        return r, path.value.decode()

    def kabsch(self, A, B, lo_a, lo_b, path):
        """Kabsch(ChainA, ChainB, LoA, LoB, Path, t, u) of the reference: returns (msd, t[3], u[3][3])."""
        t = (C.c_double * 3)()
        u = (C.c_double * 9)()

        def p(a):
            return a.ctypes.data_as(C.c_void_p)
        self.lib.ref_kabsch.restype = C.c_double
        r = self.lib.ref_kabsch(A.L, p(A.xyz[0]), p(A.xyz[1]), p(A.xyz[2]), B.L, p(B.xyz[0]), p(B.xyz[1]), p(B.xyz[2]),
                                int(lo_a), int(lo_b), path.encode(), t, u)
        return float(r), np.array(t[:]), np.array(u[:]).reshape(3, 3)

    def align_batch(self, A, B, ia, ib, nthreads):
        """A, B: SoA chain sets (objects with lens/prof/mu/xyz/selfrev numpy arrays).  Multi-threaded reference loop."""
        ia = np.ascontiguousarray(ia, np.uint32)
        ib = np.ascontiguousarray(ib, np.uint32)
        score = np.zeros(len(ia), np.float32)
        evalue = np.zeros(len(ia), np.float32)
        plen = np.zeros(len(ia), np.uint32)

        def p(a):
            return None if a is None else a.ctypes.data_as(C.c_void_p)

        self.lib.ref_align_batch(int(nthreads), A.n, p(A.lens), p(A.prof), p(A.mu), p(A.xyz), p(A.selfrev),
                                 B.n, p(B.lens), p(B.prof), p(B.mu), p(B.xyz), p(B.selfrev),
                                 C.c_uint64(len(ia)), p(ia), p(ib), p(score), p(evalue), p(plen))
        return score, evalue, plen

    def align_pair_tsv(self, A, B, columns, up, use_mu=True):
        """The reference's own TSV line (DSSAligner::ToTsv) for the pair, '' when there is no alignment."""
        def p(a):
            return None if a is None else a.ctypes.data_as(C.c_void_p)

        out = C.create_string_buffer(1 << 16)
        muA = A.mu if use_mu else None
        muB = B.mu if use_mu else None
        kA = A.kmers if use_mu else None
        kB = B.kmers if use_mu else None
        self.lib.ref_align_pair_tsv(
            A.L, (A.label or "A").encode(), A.seq, p(A.prof), p(muA), p(kA), 0 if kA is None else len(kA),
            p(A.xyz[0]), p(A.xyz[1]), p(A.xyz[2]), C.c_float(A.selfrev),
            B.L, (B.label or "B").encode(), B.seq, p(B.prof), p(muB), p(kB), 0 if kB is None else len(kB),
            p(B.xyz[0]), p(B.xyz[1]), p(B.xyz[2]), C.c_float(B.selfrev),
            columns.encode(), int(up), out, 1 << 16)
        return out.value.decode().rstrip("\n")

    def mu_score(self, a, b):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        return float(self.lib.ref_mu_score(len(a), a.ctypes.data_as(C.c_void_p), len(b), b.ctypes.data_as(C.c_void_p)))

    def mu_gapless(self, a, b):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        pa, pb = a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)
        return (float(np.float32(self.lib.ref_gapless_profb(len(a), pa, len(b), pb))),
                int(self.lib.ref_gapless_int(len(a), pa, len(b), pb)))

    def parasail_sw(self, a, b, open_=2, ext=1):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        sat = C.c_int()
        s = self.lib.ref_parasail_sw(len(a), a.ctypes.data_as(C.c_void_p), len(b), b.ctypes.data_as(C.c_void_p),
                                     open_, ext, C.byref(sat))
        return s, sat.value

    def swfast(self, S, open_, ext):
        S = np.ascontiguousarray(S, np.float32)
        LA, LB = S.shape
        lo_a, lo_b = C.c_uint32(), C.c_uint32()
        path = C.create_string_buffer(LA + LB + 2)
        s = self.lib.ref_swfast(S.ctypes.data_as(C.c_void_p), LA, LB, C.c_float(open_), C.c_float(ext),
                                C.byref(lo_a), C.byref(lo_b), path, LA + LB + 2)
        return float(np.float32(s)), lo_a.value, lo_b.value, path.value.decode()

    def lddt(self, A, B, posA, posB):
        posA = np.ascontiguousarray(posA, np.uint32)
        posB = np.ascontiguousarray(posB, np.uint32)

        def p(a):
            return a.ctypes.data_as(C.c_void_p)

        return float(np.float32(self.lib.ref_lddt(A.L, p(A.xyz[0]), p(A.xyz[1]), p(A.xyz[2]), B.L, p(B.xyz[0]),
                                                  p(B.xyz[1]), p(B.xyz[2]), p(posA), p(posB), len(posA))))

    def statsig(self, ts):
        pv, ev, q = C.c_double(), C.c_double(), C.c_double()
        self.lib.ref_statsig(C.c_double(ts), C.byref(pv), C.byref(ev), C.byref(q))
        return pv.value, ev.value, q.value
