"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/nvtt_b200.h
declares, and refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_and_library_agree(nvtt):
    hdr = open(os.path.join(ROOT, "include", "nvtt_b200.h")).read()
    declared = sorted(set(re.findall(r"NVTTB_API[^;(]*?\b(nvttb_\w+)\s*\(", hdr)))
    assert declared, "no declarations found"
    L = nvtt.lib()
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, "declared in the header but not exported: %s" % missing
    assert sorted(nvtt.EXPORTS) == declared


def test_sizes_and_mip_counts(nvtt):
    L = nvtt.lib()
    assert L.nvttb_level_size(nvtt.Format_BC1, 2048, 2048) == 512 * 512 * 8
    assert L.nvttb_level_size(nvtt.Format_BC3, 5, 3) == 2 * 1 * 16
    assert L.nvttb_level_size(nvtt.Format_BC4, 1, 1) == 8
    d = nvtt.make_process_desc(nvtt.InputFormat_BGRA_8UB, 2048, 2048, nvtt.Format_BC3, nvtt.Quality_Normal)
    assert L.nvttb_process_mip_count(d) == 12
    d = nvtt.make_process_desc(nvtt.InputFormat_BGRA_8UB, 37, 22, nvtt.Format_BC3, nvtt.Quality_Normal)
    assert L.nvttb_process_mip_count(d) == 6  # 37x22,18x11,9x5,4x2,2x1,1x1
    sizes = [(37, 22), (18, 11), (9, 5), (4, 2), (2, 1), (1, 1)]
    assert L.nvttb_process_output_size(d) == sum(((w + 3) // 4) * ((h + 3) // 4) * 16 for w, h in sizes)
    d = nvtt.make_process_desc(nvtt.InputFormat_BGRA_8UB, 256, 256, nvtt.Format_BC3, nvtt.Quality_Normal, max_level=3)
    assert L.nvttb_process_mip_count(d) == 3
    d = nvtt.make_process_desc(nvtt.InputFormat_BGRA_8UB, 256, 256, nvtt.Format_BC3, nvtt.Quality_Normal, mipmaps=False)
    assert L.nvttb_process_mip_count(d) == 1


def test_no_cpu_fallback(nvtt):
    """Without a device the context cannot be created; nothing in the product computes on the CPU."""
    if nvtt.lib().nvttb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(nvtt.NvttbError) as e:
        nvtt.Context(0)
    assert e.value.code == 4  # Error_CudaError
