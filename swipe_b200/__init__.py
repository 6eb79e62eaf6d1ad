"""swipe_b200 -- B200 (sm_100a) score-only Smith-Waterman database scan behind a C ABI.

The product is the shared library swipe_b200/csrc/libswipe_b200.so (include/swipe_b200.h);
this package is the thin host-side binding used by the tests and bench.py.  There is no CPU
implementation here: every compute call goes to the CUDA library and fails loudly without it.
"""
from .api import (Database, Scoring, SwbError, load_library, topk_merge, device_count,  # noqa: F401
                  HostBuffer, BlastDB, align, hits_merge, set_cache_limit, alu_peak)
from .scoring import (blosum62, nucleotide_matrix, parse_matrix, matrix_limits,  # noqa: F401
                      encode_protein, encode_nucleotide, SYM_AA, SYM_NT16)
