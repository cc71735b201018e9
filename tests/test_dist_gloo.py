"""CPU, world_size 2 over gloo: the row-sharding arithmetic of the multi-GPU path.
Each rank evaluates the rows of its molecule block (engine.shard_bounds, the rule csrc/comm.cu
uses) with the CPU oracle standing in for the kernels, energies are all-reduced, gradient row
slices all-gathered, and the result must equal the unsharded evaluation."""
import os
import socket

import numpy as np
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
