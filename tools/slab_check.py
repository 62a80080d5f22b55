"""torchrun entry (>= 2 GPUs): the slab-decomposed step (peer stores and NCCL halos) must
reproduce the single-GPU fused kernel BIT FOR BIT on the same initial condition.
Prints SLAB_CHECK_OK from rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lambrex_b200 import lbx                                    # noqa: E402
from lambrex_b200.slab import SlabSim, shear_slab, pulse_slab   # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    lbx.init(local)
    nx, ny, nz, tau, steps = 64, 48, 50, 0.1, 25
    ok = True
    for halo in ("p2p", "p2p-nosplit", "nccl"):
        sim = SlabSim(nx, ny, nz, tau, tau, rank=rank, world=world, halo=halo.split("-")[0],
                      split=not halo.endswith("nosplit"))
        klo, khi = sim.layout.slab(rank)
        rho, u = shear_slab(nx, ny, nz, klo, khi)
        rho = rho * pulse_slab(nx, ny, nz, klo, khi)            # z-dependence crosses the slab faces
        sim.set_initial(torch.from_numpy(rho.copy()), torch.from_numpy(u))
        sim.step(steps)
        mine = sim.download_f()
        sim.barrier()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            got = np.concatenate(parts, axis=1)
            rg, ug = shear_slab(nx, ny, nz, 0, nz - 1)
            rg = rg * pulse_slab(nx, ny, nz, 0, nz - 1)
            ref = SlabSim(nx, ny, nz, tau, tau, halo="p2p")
            ref.set_initial(torch.from_numpy(rg.copy()), torch.from_numpy(ug))
            ref.step(steps)
            want = ref.download_f()
            ref.close()
            same = np.array_equal(got, want)
            print("slab_check halo=%s world=%d bit-equal=%s max|diff|=%.3e" %
                  (halo, world, same, float(np.max(np.abs(got - want)))), flush=True)
            ok = ok and same
        sim.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    if rank == 0 and ok:
        print("SLAB_CHECK_OK", flush=True)
    dist.destroy_process_group()
    lbx.finalize()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
