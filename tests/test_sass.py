"""Static checks on the compiled objects (no GPU needed): the hot kernels really contain the Blackwell
instructions DESIGN.md claims, and the tensor-core / TMA issue paths carry no ptxas "uniformisation loops".

Background (profiles/r01_prefill_tc_notes.md): under `if (lane == 0)` ptxas cannot prove the descriptor
operands of tcgen05.mma / cp.async.bulk.tensor warp-uniform and wraps every such instruction in an
R2UR + ELECT + BRA.U.ANY loop (~100 cycles each, measured); under `if (elect_one())` (elect.sync) they are issued
back to back.  A regression to `lane == 0` would pass every parity test and only show up as lost throughput."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200", "build")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def _sass(obj):
    path = os.path.join(BUILD, obj)
    if not (os.path.exists(path) and os.path.exists(CUOBJDUMP)):
        pytest.skip("needs the built object files and cuobjdump (python __graft_entry__.py)")
    return subprocess.run([CUOBJDUMP, "-sass", path], capture_output=True, text=True, check=True).stdout


def _count(sass, needle):
    return sum(needle in line for line in sass.splitlines())


def test_prefill_kernel_uses_tcgen05_tmem_operands_and_tma():
    sass = _sass("prefill_tc.cu.o")
    assert "sm_100a" in sass
    assert _count(sass, "UTCHMMA") > 0                       # tcgen05.mma kind::f16
    assert _count(sass, "UTCHMMA tmem[") > 0                 # P.V with the A operand read from TMEM
    assert _count(sass, "UTMALDG") > 0                       # TMA tensor loads of the K/V boxes
    assert _count(sass, "UBLKCP") > 0                        # bulk copies of raw int8 units
    assert _count(sass, "BRA.U.ANY") == 0                    # no uniformisation loops anywhere in this file


def test_int8_gemm_uses_tcgen05_i8_pairs_multicast_and_no_uniformisation_loops():
    sass = _sass("gemm_i8.cu.o")
    assert _count(sass, "UTCIMMA") > 0                       # tcgen05.mma kind::i8
    assert _count(sass, ".2CTA") > 0                         # cta_group::2 MMAs / loads
    assert _count(sass, "UTMALDG.3D.MULTICAST.2CTA") > 0     # A shared between neighbouring CTA pairs
    assert _count(sass, "BRA.U.ANY") == 0


def test_decode_kernels_stream_pages_with_bulk_copies():
    sass = _sass("paged_decode.cu.o")
    assert _count(sass, "UBLKCP") > 0                        # cp.async.bulk page units (overlap kernel)
    assert _count(sass, "UTMALDG") > 0                       # TMA tensor boxes (beam-group kernel)
    assert _count(sass, "HMMA.16816") > 0                    # mma.sync m16n8k16 (beam-group kernel)
    # the producers issue their copies under elect.sync: no R2UR + BRA.U.ANY uniformisation loop around any UBLKCP /
    # UTMALDG (r01 had 40 of them in this object: every bulk copy of the streaming kernels)
    assert _count(sass, "BRA.U.ANY") == 0


def test_fp32_linear_kernel_runs_tf32_mma_with_weights_in_tensor_memory():
    sass = _sass("linear_tf32x3.cu.o")
    assert _count(sass, "UTCHMMA tmem[") >= 12               # tcgen05.mma kind::tf32, A operand (the weights) from TMEM
    assert _count(sass, "STTM") > 0                          # tcgen05.st of W / W_lo into tensor memory
    assert _count(sass, "UTMALDG.2D") > 0                    # TMA boxes of W and x
    assert _count(sass, "BRA.U.ANY") == 0
