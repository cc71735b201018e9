"""CPU, world_size 2 over gloo: the row-sharding arithmetic of the multi-GPU path.
Each rank evaluates the rows of its molecule block (engine.shard_bounds, the rule csrc/comm.cu
uses) with the CPU oracle standing in for the kernels, energies are all-reduced, gradient row
slices all-gathered, and the result must equal the unsharded evaluation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from u_pol_b200 import synth
from u_pol_b200.engine import shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    from oracle import c_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = synth.imidazole_box(11, seed=9)                       # 11 molecules: uneven shards
    D = synth.random_displacements(s, seed=10)
    lo, hi = shard_bounds(s.nmol, world, rank)
    e, g = c_oracle.energy_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u,
                                row_lo=lo * s.natoms, row_hi=hi * s.natoms)
    et = torch.as_tensor(e.copy())
    dist.all_reduce(et)                                        # C2: scalar pack
    gt = torch.as_tensor(g.reshape(-1, 3).copy())
    for src in range(world):                                   # C1: row slices, uneven -> broadcasts
        slo, shi = shard_bounds(s.nmol, world, src)
        dist.broadcast(gt[slo * s.natoms: shi * s.natoms], src=src)
    if rank == 0:
        e_full, g_full = c_oracle.energy_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u)
        q.put((abs(float(et[0]) - e_full[0]) / abs(e_full[0]), float(np.abs(gt.numpy() - g_full.reshape(-1, 3)).max())))
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_evaluation_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rel_e, max_g = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert rel_e < 1e-12 and max_g == 0.0


# ---- the same exchange driving the PRODUCT: two gloo ranks, each evaluating its molecule block on the GPU -------

def _worker_product(rank, world, port, q):
    from oracle import c_oracle                        # checker only
    from u_pol_b200.engine import UindSystem
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = synth.imidazole_box(23, seed=9)                # 23 molecules: uneven shards
    D = synth.random_displacements(s, seed=10)
    sysm = UindSystem.from_drude_system(s, device=0)   # both ranks share cuda:0 on a single-GPU box
    lo, hi = shard_bounds(s.nmol, world, rank)
    sysm.set_row_range(lo, hi)                         # upol_system_set_row_range: this rank's Drude rows only
    out = {}
    for prec in ("fp64", "mixed"):
        e, g = sysm.energy_grad(D, precision=prec)
        et = e.cpu()
        dist.all_reduce(et)                            # partial sums of the energy vector
        gt = g.cpu()
        for src in range(world):                       # row slices (other rows are exact zeros on a shard)
            slo, shi = shard_bounds(s.nmol, world, src)
            dist.broadcast(gt[slo * s.natoms: shi * s.natoms], src=src)
        u = float(et[1] + et[2] + et[3])               # E_INTER + E_INTRA + E_SELF (a shard leaves E_UIND to the caller)
        out[prec] = (u, gt.numpy())
    if rank == 0:
        e_full, g_full = c_oracle.energy_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u)
        whole = UindSystem.from_drude_system(s, device=0)
        ew, gw = whole.energy_grad(D)
        gmax = np.abs(g_full).max()
        q.put({p: (abs(out[p][0] - e_full[0]) / abs(e_full[0]), float(np.abs(out[p][1] - g_full.reshape(-1, 3)).max() / gmax))
               for p in out} | {"whole": (abs(out["fp64"][0] - float(ew[0])) / abs(float(ew[0])),
                                          float(np.abs(out["fp64"][1] - gw.cpu().numpy()).max() / gmax))})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_row_sharded_product_world2():
    """upol_system_set_row_range shards of the CUDA product, exchanged over gloo by two processes, against the
    CPU oracle and against the unsharded product (runs on a single-GPU box: both ranks use cuda:0)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_product, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res["fp64"][0] < 1e-10 and res["fp64"][1] < 1e-10, res
    assert res["mixed"][0] < 1e-5 and res["mixed"][1] < 1e-5, res
    assert res["whole"][0] < 1e-12 and res["whole"][1] < 1e-12, res
