"""GPU bring-up: the two hardware contracts the conv kernels rely on, checked in isolation.
(1) tcgen05.mma through SWIZZLE_NONE K-major descriptors with shifted start / arbitrary SBO;
(2) the 3-D TMA box load of an FT8 window."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from dfs_b200 import _probes as N  # noqa: E402   (lib/libdfs_b200_probes.so: not part of the product library)


def _bf16_bits(a):
    t = torch.from_numpy(a).to(torch.bfloat16)
    return t.view(torch.int16).numpy().view(np.uint16), t.float().numpy()


@pytest.mark.parametrize("n,k,row_shift,group_rows", [(64, 32, 0, 8), (128, 64, 0, 8), (128, 64, 3, 8), (64, 32, 1, 10),
                                                      (128, 64, 11, 10), (64, 64, 19, 18), (128, 288, 2, 10)])
def test_umma_descriptor_addressing(n, k, row_shift, group_rows):
    rng = np.random.default_rng(n + k + row_shift)
    rows_a = row_shift + 15 * group_rows + 8 + 5
    a_bits, a_f = _bf16_bits(rng.standard_normal((rows_a, k)).astype(np.float32))
    b_bits, b_f = _bf16_bits(rng.standard_normal((n, k)).astype(np.float32))
    dev = torch.device("cuda", 0)
    a_d = torch.from_numpy(a_bits.view(np.int16)).to(dev)
    b_d = torch.from_numpy(b_bits.view(np.int16)).to(dev)
    out = torch.zeros(128 * n, dtype=torch.float32, device=dev)
    N.check(N.load().dfs_probe_umma(C.c_void_p(a_d.data_ptr()), C.c_void_p(b_d.data_ptr()), rows_a, n, k, row_shift, group_rows,
                                    C.c_void_p(out.data_ptr()), None), "dfs_probe_umma")
    torch.cuda.synchronize()
    rows = np.array([row_shift + (r // 8) * group_rows + (r % 8) for r in range(128)])
    ref = a_f[rows].astype(np.float64) @ b_f.astype(np.float64).T
    got = out.cpu().numpy().reshape(128, n)
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("planes,rs,wrows,row0,col0", [(4, 162, 18, 0, 0), (4, 162, 18, 144, 16), (8, 82, 10, 72, 160), (8, 82, 10, 8, 3)])
def test_tma_window_layout(planes, rs, wrows, row0, col0):
    ncols = 200
    total = planes * ncols * rs * 8
    vals = (np.arange(total, dtype=np.int64) * 2654435761 % 65521).astype(np.uint16)
    act = vals.reshape(planes, ncols, rs, 8)
    dev = torch.device("cuda", 0)
    act_d = torch.from_numpy(act.view(np.int16)).to(dev)
    out = torch.zeros(planes * 18 * wrows * 8, dtype=torch.int16, device=dev)
    N.check(N.load().dfs_probe_tma_window(C.c_void_p(act_d.data_ptr()), planes, rs, ncols, wrows, row0, col0,
                                          C.c_void_p(out.data_ptr()), None), "dfs_probe_tma_window")
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint16).reshape(planes, 18, wrows, 8)
    ref = act[:, col0:col0 + 18, row0:row0 + wrows, :]
    assert np.array_equal(got, ref)


def test_micro_benchmarks_run():
    """The tcgen05.mma and tcgen05.ld micro-benchmarks (tools/umma_bench.py, tools/micro/tmem_ld_bench.py) return plausible cycle
    counts: an N = 128 MMA costs more than its 64-cycle floor and less than 4x that; a TMEM read moves more than 16 B/clk/SM."""
    lib = N.load()
    cyc, byt = C.c_int64(), C.c_int64()
    off = (C.c_uint32 * 32)(*([0] * 32))
    for n_acc in (1, 2):
        N.check(lib.dfs_probe_umma_bench(128, 32, 50, n_acc, off, off, 2304, 128, 2048, 128, 0, 0, C.byref(cyc), None), "umma_bench")
        per = cyc.value / (50 * 32)
        assert 60 <= per <= 256, per
    for shape, nwarps in ((2, 4), (3, 8), (6, 4)):
        N.check(lib.dfs_probe_tmem_ld_bench(shape, nwarps, 1, 200, 4, C.byref(cyc), C.byref(byt), None), "tmem_ld_bench")
        assert byt.value / cyc.value >= 16.0, (shape, nwarps, byt.value / cyc.value)
