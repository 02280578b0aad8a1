"""reseek_b200 - B200-native (sm_100a) implementation of the Reseek per-pair search hot path.

The product is the C-ABI shared library ``libreseek_b200.so`` (include/reseek_b200.h) and the C++ host
look-alikes of the reference's DSSAligner / DBSearcher in ``csrc/host``.  This Python package is only a
ctypes binding used by the tests, the benchmark and the graft entry points; it contains no compute and has
no CPU fallback: importing :mod:`reseek_b200.lib` raises if the library has not been built.
"""
from .lib import (  # noqa: F401
    ChainSet, Context, HIT_DTYPE, MODE_FAST, MODE_SENSITIVE, MODE_VERYSENSITIVE, KEEP_ALL, KEEP_HITS,
    HIT_MU_REJECTED, HIT_HAS_EVALUE, HIT_REPORTED, HIT_MKF, HIT_GLOBAL, Params, ReseekB200Error, device_count, format_aln, format_fasta2, format_tsv, kabsch, load_library, path_to_cigar,
    params_preset, prefilter_bag, version, Comm, comm_unique_id, partition_by_residues)
