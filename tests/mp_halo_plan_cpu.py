#!/usr/bin/env python3
"""World-size-2 gloo worker (CPU): each rank builds its own partition's halo plan through the C ABI's host-side query,
ships the coordinates of the nodes it would send, and checks that what arrives is exactly its own face nodes in its
own face-node order — for the trace layout of both kernel families (reference layout and aos record offsets)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import dgtd_b200 as dg
    from conftest import load_golden, product_mesh_and_kwargs
    from oracle.dgtd_oracle import HesthavenOracle

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    for name in ("box3d_p3_pec_upwind", "tfsf3d_p2_on"):
        pb, _ = load_golden(name)
        O = HesthavenOracle(pb)
        mesh, kw = product_mesh_and_kwargs(pb)
        q = lambda n, t: dg.setup_query(mesh, n, t, rank=rank, nranks=world, **kw)
        Np, Nfp = O.Np, O.Nfp
        gid = q("elem_gid", np.int32)
        send = q("send_node", np.int32)
        soff = q("wg_send_off", np.int64)
        r2d = np.argsort(q("wg_dev2ref", np.int32))
        finfo = q("finfo", np.int32).reshape(len(gid), 4, 2)
        # aos record offsets address the same nodes as the reference-layout send list
        le, node = send // Np, send % Np
        assert np.array_equal(soff, (le * Np + r2d[node]) * 6)
        xyz = O.xyz.reshape(-1, 3)
        mine = torch.from_numpy(xyz[gid[le] * Np + node].copy())
        other = torch.zeros_like(mine)
        peer = 1 - rank
        if rank == 0:
            dist.send(mine, peer); dist.recv(other, peer)
        else:
            dist.recv(other, peer); dist.send(mine, peer)
        other = other.numpy().reshape(-1, Nfp, 3)
        n = 0
        for l in range(len(gid)):
            for f in range(4):
                nb = finfo[l, f, 0]
                if nb <= -2:
                    assert np.abs(other[-2 - nb] - xyz[gid[l] * Np + O.ref.fnodes[f]]).max() < 1e-14
                    n += 1
        assert n == len(other) > 0
    dist.barrier()
    print("HALO_PLAN_OK", rank)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
