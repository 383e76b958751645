"""llm_decoder -- drop-in for the reference's pybind11 module of the same name
(src/bindings.cpp:32-35), backed by hand-written sm_100a kernels in libpa_b200.so.

Bound in the reference: CUDADecoder, INT8Decoder.  Also exposed here with the reference's
C++ signatures: KVTileCache, PageTable, AttentionCUDA, AttentionTileLauncher, int8_quant
functions and dnnl_matmul_int8.
"""
from . import _cabi  # noqa: F401
from .attention import (AttentionCUDA, AttentionTileLauncher, apply_rotary_embedding, lse_combine,  # noqa: F401
                        paged_attention_filtered, paged_decode_group, paged_decode_partial, paged_prefill)
from .int8_quant import (batch_dequantize, batch_minmax_scale, batch_quantize, compute_absmax,  # noqa: F401
                         compute_minmax_scale, dequantize_from_int8, dnnl_matmul_int8, quantize_to_int8)
from .kv_tile_cache import KVTileCache  # noqa: F401
from .page_table import PageTable  # noqa: F401
from . import sampling  # noqa: F401

from .decoders import CUDADecoder, INT8Decoder  # noqa: F401
