"""libjxl's self-contained known-answer methods for the host side of the path, restated against both
host parsers -- the oracle's (`which` = 0) and the product's csrc/host/ (`which` = 1), which share most of their text:
random field / token / permutation / section-table streams written by the oracle's writers must read back value for
value (lib/jxl/fields_test.cc, bit_reader_test.cc, ans_test.cc:27-170, :202-298, coeff_order_test.cc, toc_test.cc)."""
import ctypes

import pytest

import emul_lib


def L():
    lib = emul_lib.lib()
    for f in ("jxlb_kat_fields", "jxlb_kat_entropy", "jxlb_kat_permutation", "jxlb_kat_toc", "jxlb_kat_bit_reader"):
        getattr(lib, f).restype = ctypes.c_long
    return lib


WHICH = [pytest.param(0, id="oracle"), pytest.param(1, id="product_host")]


@pytest.mark.parametrize("which", WHICH)
def test_fields_round_trip(which):
    for seed in range(8):
        assert L().jxlb_kat_fields(ctypes.c_uint32(seed), ctypes.c_size_t(4000), which) == 0


@pytest.mark.parametrize("which", WHICH)
def test_bit_reader(which):
    for seed in range(4):
        assert L().jxlb_kat_bit_reader(ctypes.c_uint32(seed), ctypes.c_size_t(20000), which) == 0


@pytest.mark.parametrize("which", WHICH)
@pytest.mark.parametrize("mode", [pytest.param(0, id="ans"), pytest.param(1, id="prefix"), pytest.param(2, id="ans_lz77"),
                                  pytest.param(3, id="prefix_lz77")])
def test_token_streams_round_trip(which, mode):
    """ans_test.cc: RoundtripTestcase / random streams over several contexts; UintConfigs; LZ77 with and without the
    special distances of a Modular stream. The final ANS state is checked inside."""
    n = 0
    for seed in range(6):
        for num_ctx, num_clusters, count, mult, bits in [(1, 1, 50, 0, 3), (1, 1, 3000, 0, 12), (7, 3, 5000, 0, 8),
                                                         (40, 40, 20000, 97, 16), (300, 64, 20000, 0, 20),
                                                         (5, 5, 4000, 98, 31), (2, 2, 1, 0, 1)]:
            r = L().jxlb_kat_entropy(ctypes.c_uint32(seed * 131 + n), num_ctx, num_clusters, ctypes.c_size_t(count), mode, mult,
                                     bits, which)
            assert r == 0, (seed, num_ctx, num_clusters, count, mult, bits, r)
            n += 1


@pytest.mark.parametrize("which", WHICH)
def test_permutations_round_trip(which):
    """coeff_order_test.cc: Lehmer-coded permutations of every coefficient-order size (64 ... 65536 with the LLF
    coefficients skipped) read back through ReadPermutation."""
    for seed, (size, skip) in enumerate([(64, 1), (64, 1), (128, 2), (256, 4), (512, 8), (1024, 16), (2048, 32), (4096, 64),
                                         (16384, 256), (65536, 1024), (10, 0), (1, 0), (141, 0)]):
        assert L().jxlb_kat_permutation(ctypes.c_uint32(seed), ctypes.c_size_t(size), ctypes.c_size_t(skip), which) == 0


@pytest.mark.parametrize("which", WHICH)
@pytest.mark.parametrize("permuted", [0, 1])
def test_toc_round_trip(which, permuted):
    """toc_test.cc: section sizes over all four U32 ranges, with and without a permutation; offsets of the logical
    sections as lib/jxl/toc.cc:70-105 derives them."""
    for seed, entries in enumerate([1, 2, 7, 58, 141, 1000, 4100]):
        assert L().jxlb_kat_toc(ctypes.c_uint32(seed), ctypes.c_size_t(entries), permuted, which) == 0
