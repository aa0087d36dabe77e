"""Host-side logic that needs no GPU: controlDict namelist reader, grid arithmetic, workloads, fixtures."""
import os

import numpy as np
import pytest

from pixelflow_b200 import workloads as wl
from pixelflow_b200.controldict import parse_controldict

HERE = os.path.dirname(os.path.abspath(__file__))


def _deck(name):
    return np.load(os.path.join(HERE, "golden", "decks", name + ".npz"))


def test_controldict_cylinder_deck():
    cd = parse_controldict(str(_deck("cylinder")["controldict"]))
    assert cd.xnue == 0.001 and cd.width == 1.023 and cd.height == 0.511 and cd.time == 1.0
    assert cd.istep_max == 5000 and cd.istep_out == 50001 and cd.iter_max == 100 and cd.relux_factor == 1.7
    assert cd.nonslip is True and cd.threshold == 1.0e-6 and cd.thickness == 1.5
    assert cd.output_folder == "cylinder" and cd.csv_file == "data/porosity_cylinder.csv"
    assert cd.groups_seen == ["physical", "file_control", "grid_control", "porosity_control",
                              "calculation_method", "directory_control", "solver_control"]


def test_controldict_rejects_out_of_order_groups_and_unknown_keys():
    with pytest.raises(ValueError):
        parse_controldict("&solver_control\n iter_max = 3\n/\n&physical\n xnue = 1.0\n/\n")
    with pytest.raises(KeyError):
        parse_controldict("&physical\n bogus = 1.0\n/\n")
    cd = parse_controldict("&physical\n xnue = 1.5d-3, AoA = -2.0 ! comment\n/\n&calculation_method\n nonslip = .false.\n/\n")
    assert cd.xnue == 1.5e-3 and cd.AoA == -2.0 and cd.nonslip is False


def test_grid_spacing_matches_the_decks():
    cd = parse_controldict(str(_deck("room")["controldict"]))
    dx, dy, dz, dt = wl.grid_spacing(cd.width, cd.height, cd.depth, cd.time, cd.istep_max, 64, 64, 64)
    assert dx == 0.63 / 63.0 and dz == 0.63 / 63.0 and dt == 1.0 / 2000.0
    cd = parse_controldict(str(_deck("backstep")["controldict"]))
    dx, dy, _, dt = wl.grid_spacing(cd.width, cd.height, cd.depth, cd.time, cd.istep_max, 2251, 411)
    assert dx == 1.125 / 2250.0 and dy == 0.205 / 410.0 and dt == 0.1 / 2000.0


def test_deck_fixtures_have_the_surveyed_statistics():
    room = _deck("room")["porosity"]
    assert room.shape == (64, 64, 64) and abs(room.min() - 0.01798621) < 1e-12 and room.max() == 1.0
    assert abs((room >= 0.9).mean() - 0.64) < 0.02
    cyl = _deck("cylinder")["porosity"]
    assert cyl.shape == (1, 512, 1024) and cyl.max() == 1.0


def test_porous_channel_slabs_tile_the_full_field():
    m, n, l = 24, 16, 16
    full = wl.porous_channel(m, n, l, pitch=8)
    assert full.shape == (l + 2, n + 2, m + 2)
    assert full.min() >= 1e-6 and full.max() <= 1.0 and (full[1:-1, 1:-1, 1:-1] < 0.5).any()
    for first, cnt in ((1, 8), (9, 8)):
        slab = wl.porous_channel(m, n, l, pitch=8, k_first=first, k_count=cnt)
        assert np.array_equal(slab[:, 1:-1, 1:-1], np.take(full, np.arange(first - 1, first + cnt + 1), axis=0)[:, 1:-1, 1:-1])
        assert np.array_equal(slab[:, 0, 1:-1], slab[:, n, 1:-1])


def test_porosity_halo_rules():
    rng = np.random.default_rng(0)
    e = np.zeros((6, 7, 8)); e[1:-1, 1:-1, 1:-1] = rng.random((4, 5, 6))
    wl.porosity_halo_3d_periodic(e)
    assert np.array_equal(e[0], e[4]) and np.array_equal(e[5], e[1]) and np.array_equal(e[:, 0], e[:, 5])
    w = np.zeros((6, 7, 8)); w[1:-1, 1:-1, 1:-1] = rng.random((4, 5, 6))
    wl.porosity_halo_3d_wall(w)
    assert np.array_equal(w[0], w[1]) and np.array_equal(w[:, :, 0], w[:, :, 1])


def test_rank_plumbing_without_a_gpu(tmp_path):
    """pf_ranks_launch / barrier / finish (pf_ranks.cu): N forked ranks, only rank 0's stdout is the log, every rank
    writes its record at its offset of a shared file, and a failing rank is noticed by all"""
    import subprocess
    from pixelflow_b200 import build
    build.build_library()
    drv = build.build_drivers()[0]
    out = tmp_path / "ranks.txt"
    r = subprocess.run([drv, "--ranks-selftest", "5", str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == "rank 0 of 5 prints\n"
    assert out.read_text() == "".join(f"rank {k:3d} ok\n" for k in range(5))
    r = subprocess.run([drv, "--ranks-selftest", "4", str(out), "2"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and r.stdout == "rank 0 of 4 prints\n"
    r = subprocess.run([drv, "--ranks-selftest", "1", str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and out.read_text() == "rank   0 ok\n"


def test_two_dimensional_cases_refuse_several_gpus(tmp_path):
    import subprocess
    from pixelflow_b200 import build
    drv = build.build_drivers()[0]
    (tmp_path / "config").mkdir()
    r = subprocess.run([drv, "--case", "ibm2_uniform_omp", "--gpus", "2"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=60)
    assert r.returncode == 2 and "2D cases run on one GPU" in r.stderr


def test_tma_z_chunk_schedule_covers_every_plane_once():
    """pixelflow_b200/slab.py::tma_schedule / tma_block_chunk mirror pf_tma_schedule() and the kernel's block decoding
    (csrc/pf_sor_tma.cu): every plane of every tile belongs to exactly one block, no block is empty, and the picks for
    the two benchmark grids are the grid sizes ncu recorded for the product (profiles/r02_final_sor_tma_ncu_raw_*.csv:
    launch__grid_size 740 and 295)."""
    from pixelflow_b200.slab import tma_block_chunk, tma_schedule
    assert tma_schedule(1024, 512, 512)[1:] == (592, 1, 2, 740)
    assert tma_schedule(256, 256, 256)[1:] == (74, 2, 7, 295)
    assert tma_schedule(1024, 512, 64)[1:] == (592, 1, 2, 740)          # one of 8 ranks: 4 waves of whole columns + halves
    for (m, n, lz, sms) in [(1024, 512, 512, 148), (256, 256, 256, 148), (1024, 512, 64, 148), (130, 36, 34, 148),
                            (70, 20, 12, 148), (20, 12, 4, 148), (300, 64, 40, 7), (512, 130, 25, 20), (64, 48, 9, 3)]:
        tiles, tA, nzA, nzB, blocks = tma_schedule(m, n, lz, sms)
        seen = {}
        for b in range(blocks):
            tile, k0, k1 = tma_block_chunk(b, lz, tA, nzA, nzB)
            assert 0 <= tile < tiles and 1 <= k0 <= k1 <= lz, (m, n, lz, b, tile, k0, k1)
            for k in range(k0, k1 + 1):
                assert (tile, k) not in seen
                seen[(tile, k)] = b
        assert len(seen) == tiles * lz
