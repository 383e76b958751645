"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the paged-decode hot path.

`oracle.cpu`  : numpy/ctypes wrappers over oracle/liboracle_cpu.so (our line-by-line
                restatement, oracle_cpu.c; each function cites the reference file:line).
`oracle.ref`  : wrappers over oracle/_ref/libref_cpu.so, the reference's OWN
                int8_quant.cpp / softmax_lut.cpp / kv_tile_cache_cpu.cpp compiled from
                /root/reference by oracle/Makefile (not present in git history).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (llm_decoder + libpa_b200.so) never does.
"""
from . import cpu, decoder_ref, ref  # noqa: F401
