"""Import-path shim for the reference's Python callers.

api/router.py:4, web/backend_router.py:2-3 and every cli/*.py import the decoders as
`from decoder.cuda_decoder import CUDADecoder` / `from decoder.int8_decoder import INT8Decoder`
(the reference never ships such Python modules: its `decoder/` holds C++ only, the pybind module is
`llm_decoder`, src/bindings.cpp:32).  Putting this directory's parent (`<package>/compat`) on
sys.path makes those imports resolve to the B200 classes without touching the callers.
"""
