"""Stream-K tail of the tcgen05 GEMM (csrc/gemm.cu): the host-visible schedule must cover every k-block of every tile
exactly once, and the owner / contributor bookkeeping (who waits for whom, how many arrivals) must be consistent.
Runs on CPU: the plan builder is the same __host__ __device__ function the kernel's prologue calls."""
import ctypes as C

import pytest

from db1_sm100 import _lib

PAIRS = 74  # 148 SMs / 2


def _plan(lib, pair, tiles, KB, R, GS):
    out = (C.c_int * 14)()
    assert lib.db1_gemm_sk_plan(pair, C.c_longlong(tiles), KB, R, GS, out) == 0
    v = list(out)
    return dict(nseg=v[0], T_dp=v[1], tile=v[2:4], kb0=v[4:6], kb1=v[6:8], role=v[8:10], peer=v[10:12], npeer=v[12:14])


# (pair-tiles, k-blocks): the DB1-1.3B shapes at B*L = 4096 rows (M/256 x N/256, K/64) plus edge cases
SHAPES = [(128, 32), (128, 64), (64, 64), (192, 64), (192, 96), (256, 32), (256, 64), (384, 32), (512, 32),
          (2080, 32), (16 * 130, 517), (1, 64), (3, 8), (73, 9), (75, 33), (147, 16), (74, 64), (148, 64), (20, 7)]


@pytest.mark.parametrize("tiles,KB", SHAPES)
def test_stream_k_schedule_covers_every_k_block_once(tiles, KB):
    lib = _lib.lib()
    R, GS = C.c_int(0), C.c_int(0)
    on = lib.db1_gemm_sk_choose(C.c_longlong(tiles), PAIRS, KB, C.byref(R), C.byref(GS))
    R, GS = R.value, GS.value
    if tiles % PAIRS == 0:
        assert on == 0
    if not on:
        assert R == 0 and GS == 0
        return
    assert 0 < R == tiles % PAIRS and 0 < GS <= PAIRS
    cover = {}      # tile -> list of (kb0, kb1, pair, role)
    arrivals = {}   # owner pair -> contributing pairs
    owners = {}
    busy = []
    for g in range(PAIRS):
        pl = _plan(lib, g, tiles, KB, R, GS)
        assert pl["T_dp"] == tiles - R
        units = 0
        n_contrib = 0
        for n in range(pl["nseg"]):
            t, a, b, role = pl["tile"][n], pl["kb0"][n], pl["kb1"][n], pl["role"][n]
            assert pl["T_dp"] <= t < tiles and 0 <= a < b <= KB
            assert role == (0 if (a == 0 and b == KB) else 2 if a == 0 else 1)
            cover.setdefault(t, []).append((a, b, g, role))
            units += b - a
            if role == 1:
                n_contrib += 1
                arrivals.setdefault(pl["peer"][n], []).append(g)
                assert n == 0, "a contribution must be the first thing its pair does"
            if role == 2:
                owners[g] = (t, pl["peer"][n], pl["npeer"][n])
        assert n_contrib <= 1, "one workspace slot per pair"
        n_dp = len(range(g, pl["T_dp"], PAIRS))
        busy.append(units + n_dp * KB)
    # every k-block of every stream-K tile exactly once, contiguous ranges in pair order
    for t in range(tiles - R, tiles):
        segs = sorted(cover[t])
        assert segs[0][0] == 0 and segs[-1][1] == KB
        for (a0, b0, g0, _), (a1, b1, g1, _) in zip(segs, segs[1:]):
            assert b0 == a1 and g1 == g0 + 1
        if len(segs) > 1:
            owner = segs[0][2]
            assert segs[0][3] == 2 and owners[owner][0] == t
            assert owners[owner][1] == owner + 1 and owners[owner][2] == len(segs) - 1
            assert arrivals[owner] == [s[2] for s in segs[1:]]
        else:
            assert segs[0][3] == 0
    assert set(arrivals) == {g for g, o in owners.items()}
    # balance: nobody works more than one data-parallel wave less ~1 split share beyond the ideal
    ideal = tiles * KB / PAIRS
    assert max(busy) <= ideal + KB * (1 - R / PAIRS) + KB / 2 + 1
    waves_dp = -(-tiles // PAIRS) * KB
    assert max(busy) < waves_dp


def test_workspace_size_is_reported():
    lib = _lib.lib()
    lib.db1_gemm_workspace_bytes.restype = C.c_longlong
    n = lib.db1_gemm_workspace_bytes()
    assert n >= 4096 + 2 * 128 * 256 * 4
