"""`.rskc` chain-set dumps: the hand-over format between the reference's DSS stage and the host C++ look-alikes.

One file = one structure-of-arrays chain set exactly as rsk_chains_host wants it (include/reseek_b200.h), plus labels and
amino-acid sequences for the hit writers:

    "RSKC" u32 version=1  u32 n  u64 total  u32 has_mu  u32 has_selfrev
    u32 len[n] | u8 prof[8][total] | u8 mu[total]? | f32 xyz[3][total] | f32 selfrev[n]? | char seq[total] | n NUL-terminated labels

Read by reseek_b200/csrc/host/host_driver.cpp.  No compute of the hot path happens here.
"""
import struct

import numpy as np

AA = "ACDEFGHIKLMNPQRSTVWY"


def default_labels(n, prefix="chain"):
    return [f"{prefix}{i}" for i in range(n)]


def seq_from_profile(prof):
    """Amino-acid characters from the AA feature plane (feature 0, letters 0..19)."""
    lut = np.frombuffer(AA.encode(), np.uint8)
    return lut[np.minimum(np.asarray(prof[0], np.uint8), 19)]


def write_rskc(path, chains, labels=None, seq=None):
    n, total = chains.n, chains.total
    labels = labels or default_labels(n)
    assert len(labels) == n
    seq = seq_from_profile(chains.prof) if seq is None else np.frombuffer(seq.encode(), np.uint8) if isinstance(seq, str) else seq
    assert len(seq) == total
    has_mu = chains.mu is not None and len(chains.mu) == total
    has_sr = chains.selfrev is not None and len(chains.selfrev) == n
    with open(path, "wb") as f:
        f.write(b"RSKC")
        f.write(struct.pack("<IIQII", 1, n, total, int(has_mu), int(has_sr)))
        f.write(np.ascontiguousarray(chains.lens, np.uint32).tobytes())
        f.write(np.ascontiguousarray(chains.prof, np.uint8).tobytes())
        if has_mu:
            f.write(np.ascontiguousarray(chains.mu, np.uint8).tobytes())
        f.write(np.ascontiguousarray(chains.xyz, np.float32).tobytes())
        if has_sr:
            f.write(np.ascontiguousarray(chains.selfrev, np.float32).tobytes())
        f.write(np.ascontiguousarray(seq, np.uint8).tobytes())
        for lab in labels:
            f.write(lab.encode() + b"\0")
    return labels, seq
