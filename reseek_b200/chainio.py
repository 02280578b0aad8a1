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


def read_rskc(path):
    """Inverse of write_rskc: returns (SynthChains-like dict of arrays, labels, seq bytes)."""
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:4] == b"RSKC"
    ver, n, total, has_mu, has_sr = struct.unpack_from("<IIQII", buf, 4)
    pos = 4 + 24
    lens = np.frombuffer(buf, np.uint32, n, pos); pos += 4 * n
    prof = np.frombuffer(buf, np.uint8, 8 * total, pos).reshape(8, total); pos += 8 * total
    mu = None
    if has_mu:
        mu = np.frombuffer(buf, np.uint8, total, pos); pos += total
    xyz = np.frombuffer(buf, np.float32, 3 * total, pos).reshape(3, total); pos += 12 * total
    selfrev = None
    if has_sr:
        selfrev = np.frombuffer(buf, np.float32, n, pos); pos += 4 * n
    seq = np.frombuffer(buf, np.uint8, total, pos); pos += total
    labels = buf[pos:].split(b"\0")[:n]
    return dict(lens=lens, prof=prof, mu=mu, xyz=xyz, selfrev=selfrev), [x.decode() for x in labels], seq


def write_bca(path, labels, seqs, xyzs):
    """The reference's binary C-alpha format (bcadata.cpp:15-58, 147-175): magic, 3 x u64 header, per chain L amino-acid
    chars + 3L uint16 integer coordinates uint16((x + 1000) * 10 + 0.5) (pdbchain.h:89), u32 lengths, NUL-terminated labels.
    seqs: list of bytes; xyzs: list of float32 [3][L] arrays."""
    n = len(labels)
    body = bytearray()
    lens = []
    for seq, xyz in zip(seqs, xyzs):
        L = len(seq)
        assert xyz.shape == (3, L)
        ic = ((xyz.astype(np.float32) + np.float32(1000)) * np.float32(10) + 0.5).astype(np.float64)
        ic = np.floor(ic).astype(np.uint16).T.reshape(-1)  # x0 y0 z0 x1 ...
        body += bytes(seq) + ic.tobytes()
        lens.append(L)
    lens_pos = 4 + 24 + len(body)
    labeldata = b"".join(lab.encode() + b"\0" for lab in labels)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", 0xBCABCA))
        f.write(struct.pack("<QQQ", n, lens_pos, len(labeldata)))
        f.write(bytes(body))
        f.write(np.asarray(lens, np.uint32).tobytes())
        f.write(labeldata)


def read_bca(path):
    """Inverse of write_bca: returns (labels, seqs [bytes], xyzs [float32 [3][L]]), coordinates decoded as
    PDBChain::ICToCoord does (pdbchain.h:89-90: float(ic / 10.0f) - 1000)."""
    with open(path, "rb") as f:
        buf = f.read()
    magic, = struct.unpack_from("<I", buf, 0)
    assert magic == 0xBCABCA, "not a .bca file"
    n, lens_pos, label_bytes = struct.unpack_from("<QQQ", buf, 4)
    lens = np.frombuffer(buf, np.uint32, n, lens_pos)
    labels = buf[lens_pos + 4 * n:lens_pos + 4 * n + label_bytes].split(b"\0")[:n]
    seqs, xyzs = [], []
    pos = 4 + 24
    for L in lens.tolist():
        seqs.append(bytes(buf[pos:pos + L]))
        ic = np.frombuffer(buf, np.uint16, 3 * L, pos + L).reshape(L, 3)
        xyzs.append(np.ascontiguousarray((ic.astype(np.float32) / np.float32(10.0) - np.float32(1000.0)).T))
        pos += L + 6 * L
    return [x.decode() for x in labels], seqs, xyzs
