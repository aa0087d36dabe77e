"""Launched by torchrun (one process per GPU): the z-slab CUDA path against the single-domain oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank runs the CPU oracle on the whole (small) domain, then the GPU library on its slab, and
compares its own planes bit for bit.  Exit code 0 = all ranks identical.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import oracle_c
    from pixelflow_b200 import Solver, comm_unique_id
    from pixelflow_b200.slab import slab_range

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    failures = []

    def new_uid():
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, src=0)
        return bytes(buf.cpu().numpy().tobytes())

    # (sor_variant, halo_transport) pairs: 0/0 = what a user gets; 6 and 3/4 = the fused kernels on z-slabs with
    # their boundary planes sent by NCCL (1) or stored straight into the neighbours over NVLink (2: barrier kernel,
    # 3: handshake inside the TMA kernel); 1 = half-sweeps
    fused_all = [(0, 0), (6, 1), (6, 2), (6, 3), (3, 2), (4, 1), (1, 0)]
    cases = [
        # case, m, n, l, slab-host?, extra, solver options
        ("ibm3_uniform", 20, 12, 4 * world, False, {}, fused_all),                          # 4 planes per rank: the minimum
        ("ibm3_uniform", 33, 9, 4 * world + 1, True, {"xlambda": 0.1, "AoA": 5.0}, [(0, 0)]),  # odd l: colour flip at the seam
        ("ibm3_uniform", 16, 11, 3 * world + world // 2, False, {"outlet_pressure": 0.2}, [(0, 0)]),  # uneven slabs, odd n
        ("ibm3_air_condition", 14, 12, 3 * world, False, {"wall": (1, 0, 0, 0, 2, 0)}, [(0, 0)]),
        ("ibm3_air_condition", 12, 10, 4 * world, True, {"wall": (0, 2, 2, 1, 1, 2)}, [(0, 0)]),
        # top OUTLET and bottom INLET: the reference reads the opposite z face there (:702 bb(i,j,1), :948 porosity(i,j,l)),
        # which on slabs is another rank's plane
        ("ibm3_air_condition", 13, 9, 3 * world + 1, False, {"wall": (2, 1, 2, 1, 1, 2)}, [(0, 0)]),
        ("ibm3_air_condition", 10, 12, 4 * world, True, {"wall": (2, 1, 0, 0, 2, 1)}, [(0, 0)]),
        # several tiles in x and y, three z-chunks per slab; auto picks the TMA kernel + peer stores here
        ("ibm3_uniform", 130, 36, 34 * world, True, {"AoA": 3.0}, [(0, 0), (6, 1), (6, 3), (3, 2)]),
        # a long solve: 120 launches back to back replayed from the graph, the ranks' only meeting is the barrier
        # kernel (6, 2) or the in-kernel handshake (6, 3); the grid is too small for the launches to stay in step by
        # themselves
        ("ibm3_uniform", 70, 20, 12 * world, False, {"iter_max": 120}, [(6, 2), (6, 3), (0, 0)]),
        # odd planes per rank (odd colour offsets), uneven slabs when world > 2
        ("ibm3_uniform", 24, 8, 5 * world + (2 if world > 2 else 0), False, {"outlet_pressure": 0.1}, [(6, 2), (3, 1), (0, 0)]),
    ]
    used = []
    for ci, (case, m, n, l, slab_host, extra, options) in enumerate(cases):
        air = case == "ibm3_air_condition"
        rng = np.random.default_rng(100 + ci)
        kw = dict(dx=0.01, dy=0.011, dz=0.009, dt=2e-4, xnue=1e-3, xlambda=0.0, iter_max=9, relux_factor=1.7,
                  inlet_velocity=1.0, outlet_pressure=0.0, AoA=0.0)
        kw.update(extra)
        P = oracle_c.make_params(m=m, n=n, l=l, **kw)
        eps = np.clip((rng.random((l, n, m)) - 0.2) / 0.6, 1e-6, 1.0)
        if air:
            eps[-1, ::2, :] = 0.95
            eps[:, 0, ::2] = 1.0
        oc = oracle_c.Oracle3D(P, air, eps)
        for name in ("u", "v", "w", "p"):
            getattr(oc, name)[...] = 0.1 * rng.standard_normal(oc.shape)
        oc.boundary()
        first, cnt = slab_range(l, rank, world)
        sl = slice(first - 1, first + cnt + 1)
        start = [a.copy() for a in (oc.u, oc.v, oc.w, oc.p)]
        nsteps = 3
        err_o = oc.step(nsteps)
        for variant, transport in options:
            tag = f"case {ci} {case} {m}x{n}x{l} variant {variant} transport {transport} rank {rank}"
            s = Solver(case, m, n, l, device=local, rank=rank, nranks=world, nccl_unique_id=new_uid(),
                       host_is_slab=slab_host, sor_variant=variant, halo_transport=transport,
                       **{k: v for k, v in kw.items()})
            assert (s.k_first, s.k_count) == (first, cnt)
            if variant:
                assert s.sor_variant == variant, (tag, s.sor_variant)
            if transport and variant != 1:
                assert s.halo_transport == transport, (tag, s.halo_transport)
            used.append((ci, variant, transport, s.sor_variant, s.halo_transport))
            pick = (lambda a: np.ascontiguousarray(a[sl])) if slab_host else (lambda a: a)
            s.set_porosity(pick(oc.e))
            s.upload(*(pick(a) for a in start))
            err_g = s.step(nsteps)
            u, v, w, p = s.download()
            k0 = 0 if rank == 0 else 1
            k1 = cnt + 1 if rank == world - 1 else cnt
            for name, a, b in (("u", u, oc.u), ("v", v, oc.v), ("w", w, oc.w), ("p", p, oc.p)):
                mine = a[k0:k1 + 1] if slab_host else a[first - 1 + k0:first - 1 + k1 + 1]
                ref = b[first - 1 + k0:first - 1 + k1 + 1]
                if not np.array_equal(mine, ref):
                    failures.append(f"{tag}: {name} differs ({int((mine != ref).sum())} values)")
            if not np.array_equal(err_g, err_o):
                failures.append(f"{tag}: p error {err_g} vs {err_o}")
            # output_force_log_3d: slab sums reduced over the ranks, against the serial single-domain sums
            fo, fg = oc.force_log(0.05), s.force_log_3d(0.05)["raw"]
            if not np.allclose(fg, fo, rtol=1e-10, atol=1e-13 * np.abs(fo[:6]).max()):
                failures.append(f"{tag}: force log {fg} vs {fo}")
            s.close()
            dist.barrier()
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    for f in failures:
        print("FAIL", f, flush=True)
    if rank == 0:
        print("multi_gpu_check: (case, asked variant, asked transport, variant in use, transport in use):", used, flush=True)
        print(f"multi_gpu_check: world={world} cases={len(cases)} runs={len(used)} failures={int(flag.item())}", flush=True)
    dist.destroy_process_group()
    return 1 if int(flag.item()) else 0


if __name__ == "__main__":
    sys.exit(main())
