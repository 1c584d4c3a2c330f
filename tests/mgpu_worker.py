"""Worker for tests/test_multi_gpu.py (launched by torchrun, one rank per GPU): the P-rank CUDA
path must reproduce the 1-rank CUDA path (itself checked against the oracle) on the same box."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from constant_ph_b200 import capi, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lrank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    bonded = len(sys.argv) > 3 and sys.argv[3] == "bonded"
    ljstates = len(sys.argv) > 3 and sys.argv[3] == "ljstates"
    ewald = len(sys.argv) > 3 and sys.argv[3] == "ewald"
    box = synth.config(2, scale=scale, shuffle=True)
    if ewald:    # lj/cut/coul/long + the reciprocal sum; structure factors all-reduced over the ranks
        import dataclasses
        box = dataclasses.replace(box, style=capi.PAIR_COUL_LONG, alpha=0.30)
    # an atom that gains an LJ site must not be driven into its neighbours: gentler motion in that mode
    params = synth.jiggle_params(box, amp=0.35 if len(sys.argv) > 3 and sys.argv[3] == "ljstates" else 0.9,
                                 period_lo=40.0, period_hi=90.0)
    grid = bench.decompose(box, world)
    loc, sublo, subhi = bench.rank_domain(box, grid, rank)
    owned = np.nonzero(np.all((box.x >= sublo) & (box.x < subhi), axis=1))[0]
    eng = capi.Engine("cph", device=lrank)
    idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idbuf.copy_(torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    eng.comm_init_nccl(world, rank, bytes(idbuf.cpu().numpy().tobytes()))
    kw = dict(bias=dict(m_lambda=2000.0))
    if bonded:   # bond / angle partners of atoms near a sub-box face are ghosts
        kw["topology"] = synth.topology(box)
    if ljstates:  # atoms with LJ end states act on the owned atoms of neighbouring ranks as ghosts
        kw["lj_typeB"] = synth.lj_end_state_types(box)
    if ewald:
        kw["kspace"] = dict(g_ewald=box.alpha, kmax=(8, 8, 8))
    capi.configure(eng, box, sublo=sublo, subhi=subhi, procgrid=grid, myloc=loc, owned=owned, **kw)
    ref = capi.configure(capi.Engine("cph", device=lrank), box, **kw) if rank == 0 else None

    worst = dict(f=0.0, lam=0.0, dudl=0.0, e=0.0)
    f_loc = np.zeros((owned.size, 3))
    f_ref = np.zeros((box.n, 3))
    for step in range(nsteps):
        x = synth.jiggle_positions(box, params, step * box.dt)
        eng.post_force(step, box.dt, x[owned], f_loc)
        # forces of all ranks, by global index
        full = torch.zeros((box.n, 3), dtype=torch.float64, device="cuda")
        full[torch.from_numpy(owned).cuda()] = torch.from_numpy(f_loc).cuda()
        dist.all_reduce(full)
        s_m, t_m = eng.get_scalars(), eng.get_sites()
        if rank == 0:
            ref.post_force(step, box.dt, x, f_ref)
            s_r, t_r = ref.get_scalars(), ref.get_sites()
            fm = full.cpu().numpy()
            worst["f"] = max(worst["f"], np.abs(fm - f_ref).max() / np.abs(f_ref).max())
            worst["lam"] = max(worst["lam"], np.abs(t_m["lambda"] - t_r["lambda"]).max())
            worst["dudl"] = max(worst["dudl"], np.abs(t_m["dudl"] - t_r["dudl"]).max() / np.abs(t_r["dudl"]).max())
            for k in ("HA", "HB", "evdwl", "ecoul", "H_lambda"):
                worst["e"] = max(worst["e"], abs(s_m[k] - s_r[k]) / abs(s_r[k]))
        if bonded:
            eb = eng.get_bonded_energy()          # collective: all ranks call it
            if rank == 0:
                worst["e"] = max(worst["e"], float(np.abs(eb - ref.get_bonded_energy()).max() / ref.get_bonded_energy().max()))
    counts = eng.get_counts()
    tot = torch.tensor([counts["nlocal"], counts["neighbors"], counts["titr_owned"], counts["builds"]],
                       dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    if rank == 0:
        c_r = ref.get_counts()
        out = dict(worst=worst, nlocal=int(tot[0]), neighbors=int(tot[1]), titr=int(tot[2]), builds=counts["builds"],
                   ref_nlocal=c_r["nlocal"], ref_neighbors=c_r["neighbors"], ref_titr=c_r["titr_owned"],
                   ref_builds=c_r["builds"], world=world, nghost=counts["nghost"], halo=eng.get_halo_mode())
        print("MGPU_RESULT " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
