"""world_size-2 (and 4) gloo test of the host-side N>1 logic on CPU: the brick decomposition used
by bench.py owns every atom exactly once, and the reduction layout of cpp:274's replacement
([HA, HB, E_vdwl, E_coul, dU/dlambda_s..., (HB-HA)_s...], owned atoms only, then all-reduce)
reproduces the global site sums."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    import bench
    from constant_ph_b200 import capi, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    box = synth.config(2, scale=0.25, shuffle=True)
    grid = bench.decompose(box, world)
    loc, sublo, subhi = bench.rank_domain(box, grid, rank)
    owned = np.all((box.x >= sublo) & (box.x < subhi), axis=1)
    cover = torch.from_numpy(owned.astype(np.int64))
    dist.all_reduce(cover)
    # replicated oracle pass; each rank reduces ONLY its owned atoms (cpp:264: i < nlocal)
    capi.load_library("orc").orc_set_threads(2)
    o = capi.configure(capi.Engine("orc"), box)
    o.pair_pass(1); o.site_reduce()
    e, phi, q = o.get_eatom(), o.get_phi(), o.get_q()
    S = box.nsites
    red = np.zeros(4 + 2 * S)
    H = (box.mask & synth.GROUP_H_BIT) != 0
    red[0] = e[owned].sum()
    red[1] = e[owned & ~H].sum()
    ecoul = 0.5 * (q * phi)[owned].sum()
    red[2] = red[0] - ecoul
    red[3] = ecoul
    idx = box.meta["tag_to_index"][box.titr_tag]
    for t, i in enumerate(idx):
        if owned[i]:
            s = box.titr_site[t]
            red[4 + s] += (box.qB[t] - box.qA[t]) * phi[i]
            if H[i]:
                red[4 + S + s] -= e[i]
    tr = torch.from_numpy(red)
    dist.all_reduce(tr)
    if rank == 0:
        sc, st = o.get_scalars(), o.get_sites()
        ok = bool((cover == 1).all())
        err = max(abs(tr[0].item() - sc["HA"]) / abs(sc["HA"]), abs(tr[1].item() - sc["HB"]) / abs(sc["HB"]),
                  abs(tr[2].item() - sc["evdwl"]) / abs(sc["evdwl"]), abs(tr[3].item() - sc["ecoul"]) / abs(sc["ecoul"]),
                  np.abs(tr[4:4 + S].numpy() - st["dudl"]).max() / np.abs(st["dudl"]).max(),
                  np.abs(tr[4 + S:].numpy() - st["hdiff"]).max() / max(np.abs(st["hdiff"]).max(), 1e-300))
        out_q.put((ok, float(err), [int(v) for v in grid]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_owned_atom_reduction_over_ranks(built, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, err, grid = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok, "decomposition does not own every atom exactly once"
    assert err < 1e-10, err
    assert int(np.prod(grid)) == world


def bonded_worker(rank, world, port, out_q):
    """SURVEY 8 f2 on P ranks: every rank evaluates the bonds and angles its OWNED atoms take part in (LAMMPS'
    newton_bond-off lists) and keeps only those atoms' energy shares; no reverse exchange.  The all-reduced
    shares must equal the oracle's once-per-term totals."""
    sys.path.insert(0, ROOT)
    import bench
    from constant_ph_b200 import capi, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    box = synth.config(2, scale=0.25, shuffle=True)
    topo = synth.topology(box)
    grid = bench.decompose(box, world)
    loc, sublo, subhi = bench.rank_domain(box, grid, rank)
    owned = np.nonzero(np.all((box.x >= sublo) & (box.x < subhi), axis=1))[0]
    L = box.boxhi - box.boxlo
    t2i = box.meta["tag_to_index"]

    def delta(a, b):
        d = box.x[a] - box.x[b]
        return d - L * np.round(d / L)

    eb = ea = 0.0
    reach = 0.0          # farthest bonded partner of an owned atom: must lie inside the ghost shell
    for m in range(topo.maxbond):
        rows = owned[topo.num_bond[owned] > m]
        j = t2i[topo.bond_atom[rows, m]]
        r = np.linalg.norm(delta(rows, j), axis=1)
        bt = topo.bond_type[rows, m]
        eb += (0.5 * topo.bond_k[bt] * (r - topo.bond_r0[bt]) ** 2).sum()
        reach = max(reach, r.max() if r.size else 0.0)
    for m in range(topo.maxangle):
        rows = owned[topo.num_angle[owned] > m]
        i1, i2, i3 = (t2i[a[rows, m]] for a in (topo.angle_atom1, topo.angle_atom2, topo.angle_atom3))
        d1, d2 = delta(i1, i2), delta(i3, i2)
        c = (d1 * d2).sum(axis=1) / np.linalg.norm(d1, axis=1) / np.linalg.norm(d2, axis=1)
        at = topo.angle_type[rows, m]
        ea += (topo.angle_k[at] * (np.arccos(np.clip(c, -1, 1)) - topo.angle_theta0[at]) ** 2 / 3.0).sum()
        for other in (i1, i2, i3):
            dd = np.linalg.norm(delta(rows, other), axis=1)
            reach = max(reach, dd.max() if dd.size else 0.0)
    tot = torch.tensor([eb, ea, 0.0], dtype=torch.float64)
    dist.all_reduce(tot)
    far = torch.tensor([reach], dtype=torch.float64)
    dist.all_reduce(far, op=dist.ReduceOp.MAX)
    if rank == 0:
        capi.load_library("orc").orc_set_threads(2)
        o = capi.configure(capi.Engine("orc"), box, topology=topo)
        o.pair_pass(1)
        ref = o.get_bonded_energy()
        err = max(abs(tot[0].item() - ref[0]) / ref[0], abs(tot[1].item() - ref[1]) / ref[1])
        out_q.put((float(err), float(far[0]), float(box.cut_coul + box.skin)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_bonded_owned_shares_over_ranks(built, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29720 + world
    procs = [ctx.Process(target=bonded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    err, reach, rlist = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert err < 1e-12, err
    assert reach < rlist          # every partner is an owned atom or a ghost the halo already carries
