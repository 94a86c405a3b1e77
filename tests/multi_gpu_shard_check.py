"""Run under torchrun (one rank per GPU) by tests/test_gpu_multi.py: ONE map spatially sharded over the ranks (ksg_config.shard_rank /
shard_count): rank 0 owns the camera stream and broadcasts every frame with NCCL, every rank integrates it and applies only
the tiles it owns; the per-rank exports are assembled with ksg_owner_mask and must equal the map of an unsharded integrator bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kimera_semantics_b200 import synth  # noqa: E402
from kimera_semantics_b200.capi import Integrator, KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED, merge_shard_exports  # noqa: E402
from parity_utils import compare_maps, make_config  # noqa: E402


def main():
    out_dir = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, C, NF = 320, 240, 21, 3
    cam = synth.make_camera(W, H)
    report = []
    for itype, name in ((KSG_INTEGRATOR_FAST, "fast"), (KSG_INTEGRATOR_MERGED, "merged")):
        cfg = make_config(itype, 0.05, C, max_points=W * H, max_updates=16 << 20, device=local, shard_rank=rank, shard_count=world)
        integ = Integrator(cfg)
        ref = None
        if rank == 0:
            ref = Integrator(make_config(itype, 0.05, C, max_points=W * H, max_updates=16 << 20, device=local))
        buf_d = torch.empty((H, W), dtype=torch.float32, device="cuda")
        buf_l = torch.empty((H, W), dtype=torch.uint8, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        for f in range(NF):
            depth, label, T = synth.frame(cam, f, C)      # poses are known to every rank; pixels come from rank 0 only
            if rank == 0:
                buf_d.copy_(torch.from_numpy(depth))
                buf_l.copy_(torch.from_numpy(label))
                ref.integrate_depth(T, depth, label, cam.K)
            dist.broadcast(buf_d, 0)
            dist.broadcast(buf_l, 0)
            torch.cuda.synchronize()
            integ.integrate_depth_device(T, buf_d.data_ptr(), buf_l.data_ptr(), W, H, cam.K, stream, want_stats=True)
        exp = integ.export()
        gathered = [None] * world
        dist.gather_object(exp, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            merged = merge_shard_exports(gathered, cfg.voxels_per_side)
            rep = compare_maps(merged, ref.export())
            ok = (rep["same_blocks"] == 1.0 and rep["label_mismatch"] == 0 and rep["tsdf_distance_bit_mismatch"] == 0
                  and rep["tsdf_weight_bit_mismatch"] == 0 and rep["sem_priors_bit_mismatch"] == 0 and rep["tsdf_rgba_mismatch"] == 0)
            report.append(f"{name}: ranks={world} blocks={int(rep['blocks_a'])} observed={int(rep.get('observed_voxels', 0))} "
                          f"bit mismatches d/w/p={int(rep.get('tsdf_distance_bit_mismatch', -1))}/{int(rep.get('tsdf_weight_bit_mismatch', -1))}/"
                          f"{int(rep.get('sem_priors_bit_mismatch', -1))} -> {'OK' if ok else 'FAIL'}")
            ref.close()
        integ.close()
    if rank == 0:
        with open(os.path.join(out_dir, "shard_check.txt"), "w") as fh:
            fh.write("\n".join(report) + "\n")
        print("\n".join(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
