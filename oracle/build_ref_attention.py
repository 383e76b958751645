#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- build the reference's OWN cpu_paged_attention_forward into oracle/_ref.

attention_cpu/cpu_attention_kernel.cpp does not compile as shipped: the struct
member / local `int T` shadows the template parameter `T`, `Vec` is used before any
using-declaration and with a 2-argument load/store that vec_cpu.hpp does not have, and
`kv_cache->get` is called with a fourth 'k'/'v' argument KVTileCacheCPU does not take.  This
script applies the MINIMAL identifier-level edits below to a temporary copy (never written into
the repository: the copy lives in a tempfile.TemporaryDirectory and is deleted after the compile),
compiles it in front of oracle/ref_attention_shim.hpp and oracle/ref_attention_entry.cpp, and
links oracle/_ref/libref_attn.so.  Every loop, constant and arithmetic statement of
cpu_attention_kernel.cpp:36-129 is compiled verbatim; the decoder's header-only LayerNorm / MLP /
TokenEmbedding (decoder/layer_norm.hpp, mlp.hpp, token_embedding.hpp) compile unmodified.

Edits (each asserted to match exactly the expected number of times, so a changed reference fails
loudly instead of silently producing something else):
  1. template parameter `T` -> `Tq` wherever it names the ELEMENT TYPE (template headers,
     `const T*`, `T* out`, `<T>` arguments); the int `T` (sequence length) keeps its name;
  2. `using refshim::Vec;` after the includes, the late `using cpuvec::Vec;` removed;
  3. `KVTileCacheCPU<T>* kv_cache` -> `refshim::KVTileStore4<Tq>* kv_cache` (4-argument get).
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REF", "/root/reference")
OUT = os.path.join(HERE, "_ref", "libref_attn.so")
CXX = os.environ.get("ORC_CXX", "/usr/bin/g++")
# the flags of oracle/Makefile's `ref` target (the reference ships no usable flags of its own)
CXXFLAGS = ["-std=c++17", "-O2", "-fopenmp", "-fPIC", "-fpermissive", "-w"]


def _sub(text, pattern, repl, count, what):
    new, n = re.subn(pattern, repl, text)
    if n != count:
        raise SystemExit(f"build_ref_attention: edit '{what}' matched {n} times, expected {count} "
                         f"(the reference source changed?)")
    return new


def patched_sources():
    with open(os.path.join(REF, "attention_cpu", "cpu_attention_kernel.hpp")) as f:
        hpp = f.read()
    with open(os.path.join(REF, "attention_cpu", "cpu_attention_kernel.cpp")) as f:
        cpp = f.read()
    hpp = _sub(hpp, r"template <typename T>", "template <typename Tq>", 3, "hpp template headers")
    hpp = _sub(hpp, r"const T\* q\b", "const Tq* q", 1, "hpp q pointer")
    hpp = _sub(hpp, r"\bT\* out\b", "Tq* out", 1, "hpp out pointer")
    hpp = _sub(hpp, r"KVTileCacheCPU<T>\* kv_cache", "refshim::KVTileStore4<Tq>* kv_cache", 1, "hpp 4-arg store")
    hpp = _sub(hpp, r"CPUAttention(Input|Output)<T>", r"CPUAttention\1<Tq>", 2, "hpp function signature")
    cpp = _sub(cpp, r"template <typename T>", "template <typename Tq>", 2, "cpp template headers")
    cpp = _sub(cpp, r"(#include <omp\.h>\n)", r"\1using refshim::Vec;\n", 1, "cpp early using")
    cpp = _sub(cpp, r"[ \t]*using cpuvec::Vec;\n", "", 1, "cpp late using")
    cpp = _sub(cpp, r"CPUAttention(Input|Output)<T>", r"CPUAttention\1<Tq>", 2, "cpp function signature")
    cpp = _sub(cpp, r"const T\* (q_ptr|k_tile|v_tile)\b", r"const Tq* \1", 3, "cpp element pointers")
    cpp = _sub(cpp, r"apply_rotary_embedding_tile<T>\(", "apply_rotary_embedding_tile<Tq>(", 1, "cpp rope call")
    return hpp, cpp


def build(force=False):
    srcs = [os.path.join(REF, "attention_cpu", f) for f in ("cpu_attention_kernel.hpp", "cpu_attention_kernel.cpp",
                                                           "softmax_lut.cpp")]
    srcs += [os.path.join(REF, "kv_cache", "kv_tile_cache_cpu.cpp"),
             os.path.join(REF, "decoder", "layer_norm.hpp"), os.path.join(REF, "decoder", "mlp.hpp"),
             os.path.join(REF, "decoder", "token_embedding.hpp"),
             os.path.join(HERE, "ref_attention_shim.hpp"), os.path.join(HERE, "ref_attention_entry.cpp"), __file__]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in srcs):
        return OUT
    hpp, cpp = patched_sources()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="ref_attn_") as tmp:
        os.makedirs(os.path.join(tmp, "attention_cpu"))
        with open(os.path.join(tmp, "attention_cpu", "cpu_attention_kernel.hpp"), "w") as f:
            f.write(hpp)
        patched_cpp = os.path.join(tmp, "attention_cpu", "cpu_attention_kernel.cpp")
        with open(patched_cpp, "w") as f:
            f.write(cpp)
        inc = ["-I", os.path.join(tmp, "attention_cpu"), "-I", REF, "-I", os.path.join(REF, "attention_cpu"),
               "-I", HERE]
        objs = []
        units = [(patched_cpp, ["-include", os.path.join(HERE, "ref_attention_shim.hpp")]),
                 (os.path.join(REF, "attention_cpu", "softmax_lut.cpp"), []),
                 (os.path.join(REF, "kv_cache", "kv_tile_cache_cpu.cpp"), []),
                 (os.path.join(HERE, "ref_attention_entry.cpp"), ["-include", os.path.join(HERE, "ref_attention_shim.hpp")])]
        for i, (src, extra) in enumerate(units):
            obj = os.path.join(tmp, f"u{i}.o")
            subprocess.check_call([CXX, *CXXFLAGS, *inc, *extra, "-c", src, "-o", obj])
            objs.append(obj)
        subprocess.check_call([CXX, "-shared", "-fopenmp", "-o", OUT, *objs])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
