"""The GPU tools and bench legs, run on the CPU against a FAKE libpixelflow_gpu.so.

Round 1 lost its deck probe and its variant-7 run on the driver's box to `s.sor_variant()` -- a property called like a
method -- in code that only ever executes on a GPU.  Here the real `pixelflow_b200.api.Solver` class (properties,
argument marshalling, shapes) runs over a stand-in for the ctypes library whose entry points succeed and compute
nothing, so attribute misuse, bad keyword arguments and JSON-serialisation slips in tools/*.py and in bench.py's
GPU-side helpers surface in `-m "not gpu"`.  Nothing here says anything about numerics.
"""
import ctypes as C
import importlib.util
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Fn:
    def __init__(self, lib, name):
        self.lib, self.name, self.argtypes, self.restype = lib, name, None, C.c_int

    def __call__(self, *a):
        lib, nm = self.lib, self.name
        lib.calls.append(nm)
        if nm == "pf_abi_version":
            return 2
        if nm == "pf_config_init":
            cfg = a[0]._obj
            cfg.struct_size = C.sizeof(type(cfg))
            cfg.nranks, cfg.device, cfg.use_graph = 1, -1, 1
            return None
        if nm == "pf_create":
            cfg = a[1]._obj
            lib.cfg = {f: getattr(cfg, f) for f in ("m", "n", "l", "solver_case", "sor_variant", "nranks", "rank")}
            a[0]._obj.value = 0x1000
            return 0
        if nm == "pf_local_slab":
            base, rem = divmod(lib.cfg["l"], lib.cfg["nranks"])
            r = lib.cfg["rank"]
            a[1]._obj.value = r * base + min(r, rem) + 1
            a[2]._obj.value = base + (1 if r < rem else 0)
            return 0
        if nm == "pf_last_error":
            return b""
        if nm == "pf_get_sor_variant":
            return lib.cfg["sor_variant"] or 1
        if nm == "pf_get_halo_transport":
            return 0
        if nm == "pf_last_timing":
            a[1]._obj.value, a[2]._obj.value, a[3]._obj.value = 2.0, 1.0, 7
            return 0
        if nm == "pf_stream":
            return 0
        return 0


class FakeLib:
    def __init__(self):
        self.calls, self.cfg, self._fns = [], {}, {}

    def __getattr__(self, name):
        if name.startswith("pf_"):
            return self._fns.setdefault(name, _Fn(self, name))
        raise AttributeError(name)


@pytest.fixture()
def fake(monkeypatch):
    from pixelflow_b200 import api
    lib = FakeLib()
    monkeypatch.setattr(api, "load_library", lambda: lib)
    return lib


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_solver_properties_are_properties(fake):
    from pixelflow_b200 import Solver
    s = Solver("ibm3_uniform", 8, 6, 4, dx=1, dy=1, dz=1, dt=1, xnue=1e-3, sor_variant=6)
    assert s.sor_variant == 6 and s.halo_transport == 0
    assert isinstance(type(s).sor_variant, property) and isinstance(type(s).halo_transport, property)
    assert s.last_timing() == {"ms_total": 2.0, "ms_sor": 1.0, "launches": 7}
    assert s.step(3).shape == (3,)
    s.close()


@pytest.mark.parametrize("variant", [0, 7])
def test_decks_probe_runs_end_to_end(fake, monkeypatch, capsys, variant):
    mod = _load(os.path.join(ROOT, "tools", "decks_probe.py"), "decks_probe_under_test")
    monkeypatch.setattr(sys, "argv", ["decks_probe.py", "--sor-variant", str(variant), "--steps", "2"])
    mod.main()
    rows = [json.loads(ln) for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert [r["deck"] for r in rows] == ["cylinder", "backstep", "room"]
    for r in rows:
        assert r["sor_variant"] == (variant or 1)
        assert r["bit_identical_to_reference_after_3_steps"] is False   # the fake computes nothing
        assert r["ms_per_step"] == 1.0


def test_bench_decks_tool_runs_end_to_end(fake, monkeypatch, capsys):
    mod = _load(os.path.join(ROOT, "tools", "bench_decks.py"), "bench_decks_under_test")
    monkeypatch.setattr(sys, "argv", ["bench_decks.py", "--steps", "1", "--cpu-steps", "1"])
    monkeypatch.setenv("OMP_NUM_THREADS", "4")
    mod.main()
    rows = [json.loads(ln) for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert [r["deck"] for r in rows] == ["cylinder", "backstep", "room"]
    assert all(r["sor_variant"] == 1 and r["gpu_ms_per_step"] == 2.0 for r in rows)


def test_bench_field_digest_is_decomposition_independent():
    """bench.py's `parity` key: per-plane SHA-256 digests, combined in global plane order -- the same bytes whether
    one rank holds the whole grid or N ranks hold z-slabs of it"""
    sys.path.insert(0, ROOT)
    import bench
    rng = np.random.default_rng(5)
    l, n, m = 12, 5, 7
    full = [rng.standard_normal((l + 2, n + 2, m + 2)) for _ in range(4)]
    one = bench.combine_plane_digests([bench.plane_digests(full, 1, l, 0, 1)], l)
    for nranks in (2, 3, 4):
        parts = []
        base, rem = divmod(l, nranks)
        for r in range(nranks):
            k0 = r * base + min(r, rem) + 1
            kc = base + (1 if r < rem else 0)
            slab = [a[k0 - 1:k0 + kc + 1].copy() for a in full]
            for a in slab:                      # inner ghost planes may hold anything: they are not hashed
                if r > 0:
                    a[0] = -1.0
                if r < nranks - 1:
                    a[-1] = -2.0
            parts.append(bench.plane_digests(slab, k0, kc, r, nranks))
        assert bench.combine_plane_digests(parts, l) == one
    full[3][4, 2, 3] = np.nextafter(full[3][4, 2, 3], np.inf)   # one ulp anywhere changes the hash
    assert bench.combine_plane_digests([bench.plane_digests(full, 1, l, 0, 1)], l) != one
