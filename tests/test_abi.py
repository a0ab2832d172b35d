"""CPU suite, part 2: the C-ABI boundary — the library builds for sm_100a, loads, exports every
symbol include/sbir_b200.h declares, the ctypes prototypes cover the header one to one, and the
argument checks that return before touching the device behave (no compute without a GPU)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "sbir_b200.h"


def _declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(sbir_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(sbir_lib):
    from art_sbir_b200 import _binding
    names = _declared_symbols()
    assert len(names) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", str(_binding.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (sbir_[a-z0-9_]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    # nothing torch-typed or C++-mangled is part of the contract
    assert all(hasattr(sbir_lib, n) for n in names)


def test_binding_covers_header_exactly(sbir_lib):
    from art_sbir_b200 import _binding
    assert sorted(_binding.PROTOTYPES) == _declared_symbols()
    assert sbir_lib.sbir_abi_version() == _binding.ABI_VERSION == 2


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "sbir_b200.h"\nint main(void){return SBIR_OK + SBIR_F32 + SBIR_EUCLIDEAN;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-c", str(src), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_library_is_sm100a_with_tcgen05_and_tma(sbir_lib):
    from art_sbir_b200 import _binding
    elf = subprocess.run(["cuobjdump", "-lelf", str(_binding.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run(["cuobjdump", "-sass", str(_binding.LIB_PATH)], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCQMMA" in sass      # tcgen05.mma
    assert "UTMALDG" in sass                            # TMA tensor loads
    assert "LDTM" in sass                               # tcgen05.ld


def test_status_strings_and_argument_checks(sbir_lib):
    lib = sbir_lib
    assert lib.sbir_status_string(0) == b"ok"
    assert b"invalid" in lib.sbir_status_string(1)
    assert b"workspace" in lib.sbir_status_string(4)
    # bad enum / sizes are rejected before any CUDA call
    assert lib.sbir_l2_normalize(None, None, 4, 8, 7, 1e-8, None) == 1
    assert lib.sbir_pairwise_distance(None, 3, None, 2, 8, 0, 0, None, None) == 1      # 3 vs 2 rows do not broadcast
    assert lib.sbir_pairwise_topk(None, 4, None, None, 4, 8, 0, 0, 0, 0, None, None, None, None, None, None, 0, None) == 1  # k = 0
    assert lib.sbir_pairwise_topk(None, 4, None, None, 4, 8, 0, 0, 500, 0, None, None, None, None, None, None, 0, None) == 2  # k too large
    assert lib.sbir_pairwise_topk(None, 4, None, None, 4, 8, 0, 5, 10, 0, None, None, None, None, None, None, 0, None) == 1  # bad metric
    assert lib.sbir_triplet_margin_loss(None, None, None, 4, 8, 0.2, 0, None, None, None, None, None, None) == 1
    assert lib.sbir_pairwise_topk_workspace_bytes(1000, 10000, 2048, 10, 0, 0, 1) > 0
    assert lib.sbir_pairwise_topk_workspace_bytes(1000, 10000, 2048, 1000, 0, 0, 1) == 0
    # host-buffer entry points: argument checks come before any device work
    assert lib.sbir_retrieve_host(None, 4, None, 4, 8, 0, 0, 10, None, None, None, None, None) == 1            # NULL buffers
    assert lib.sbir_retrieve_host(None, 0, None, 4, 8, 0, 0, 10, None, None, None, None, None) == 1            # no queries
    assert lib.sbir_retrieve_host_shard(None, 4, None, 4, 8, 0, 0, 10, 0, None, None, None, None, None, None, None) == 1
    assert lib.sbir_retrieve_host_shard(None, 4, None, 4, 8, 9, 0, 10, 0, None, None, None, None, None, None, None) == 1  # bad dtype
    assert lib.sbir_debug_k1_diag(None, 8) == 1
    # gallery append: the block must fit the preallocated matrix; bad dtypes / NULL buffers are rejected up front
    assert lib.sbir_gallery_append(None, 0, 4, 8, None, 0, 3, 0, None, 0, None) == 1        # 4 rows into a 3-row gallery
    assert lib.sbir_gallery_append(None, 0, 2, 8, None, 0, 3, 2, None, 0, None) == 1        # rows [2, 4) of 3
    assert lib.sbir_gallery_append(None, 7, 2, 8, None, 0, 3, 0, None, 0, None) == 1        # bad block dtype
    assert lib.sbir_gallery_append(None, 0, 2, 8, None, 0, 3, 0, None, 0, None) == 1        # NULL buffers
    assert lib.sbir_gallery_append(None, 0, 0, 8, None, 1, 3, 3, None, 1, None) == 0        # empty block: no-op
    # tuning / test switches: set through the ABI, never read from the environment on the launch path
    assert lib.sbir_debug_set_option(b"k1_chunk_mb", 1) == 0 and lib.sbir_debug_set_option(b"reset", 0) == 0
    assert lib.sbir_debug_set_option(b"no_such_option", 1) == 1 and lib.sbir_debug_set_option(None, 1) == 1
    assert lib.sbir_debug_diag_build() == 0                                                   # product build: diagnostics compiled out
    # empty problems are a no-op success
    assert lib.sbir_l2_normalize(None, None, 0, 8, 0, 1e-8, None) == 0
    assert lib.sbir_pairwise_topk(None, 0, None, None, 4, 8, 0, 0, 10, 0, None, None, None, None, None, None, 0, None) == 0


def test_gather_rows_host(sbir_lib):
    """Host utility of the sharded host path (no device involved): rows picked by index, zero rows for indices
    outside the shard, any thread count."""
    import torch
    src = torch.arange(40 * 1000, dtype=torch.float32).reshape(1000, 40)
    index = torch.tensor([5, 999, -1, 1000, 0, 5] * 5000, dtype=torch.int64)
    for threads in (0, 1, 3):
        dst = torch.full((index.numel(), 40), -7.0)
        assert sbir_lib.sbir_gather_rows_host(src.data_ptr(), 1000, 160, index.data_ptr(), index.numel(), dst.data_ptr(), threads) == 0
        ok = (index >= 0) & (index < 1000)
        assert torch.equal(dst[ok], src[index[ok]]) and (dst[~ok] == 0).all()
    assert sbir_lib.sbir_gather_rows_host(None, 10, 160, None, 4, None, 0) == 1
    assert sbir_lib.sbir_gather_rows_host(None, 10, 160, None, 0, None, 0) == 0


def test_launch_path_reads_no_environment_variables():
    """tuning switches must not be getenv() calls on the launch path."""
    for src in (ROOT / "art_sbir_b200" / "csrc").glob("*"):
        assert "getenv" not in src.read_text(), src


def test_debug_options_change_the_plan_and_reset(sbir_lib):
    out = (ctypes.c_int32 * 13)()
    from art_sbir_b200 import _binding
    try:
        assert sbir_lib.sbir_debug_plan(100_000, 10_000_000, 512, 10, 1, 148, out) == 0
        auto_chunks = out[6]
        _binding.set_debug_option("k1_chunk_mb", 12)
        assert sbir_lib.sbir_debug_plan(100_000, 10_000_000, 512, 10, 1, 148, out) == 0
        assert out[6] > auto_chunks                                   # smaller chunk steps -> more of them
        _binding.set_debug_option("k1_pair", 2)
        assert sbir_lib.sbir_debug_plan(12_500, 75_000, 1024, 10, 0, 148, out) == 0 and out[10] == 2
    finally:
        _binding.set_debug_option("reset")
    assert sbir_lib.sbir_debug_plan(100_000, 10_000_000, 512, 10, 1, 148, out) == 0 and out[6] == auto_chunks
    assert sbir_lib.sbir_debug_plan(12_500, 75_000, 1024, 10, 0, 148, out) == 0 and out[10] == 1
    # bf16 tiles of rows >= 4 KB (2048-d fp32 embeddings selected on their bf16 copies) take CTA pairs on their own
    assert sbir_lib.sbir_debug_plan(12_500, 75_000, 2048, 10, 0, 148, out) == 0 and out[10] == 2 and out[12] == 1


def test_product_refuses_cpu_tensors_and_never_imports_the_oracle():
    import torch
    from art_sbir_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.pairwise_topk(torch.randn(4, 8), torch.randn(9, 8), 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.triplet_margin_loss(torch.randn(4, 8), torch.randn(4, 8), torch.randn(4, 8))
    for py in (ROOT / "art_sbir_b200").glob("*.py"):
        assert "oracle" not in py.read_text().replace("oracle.synthetic_embeddings", ""), py
