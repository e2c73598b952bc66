"""Launched with torchrun on N GPUs: stitches one strip panorama sharded by column strip (NCCL halo exchange) and
checks the assembled strips against the single-GPU pipeline on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py [rows cols per_rank [grid_rows]]
grid_rows > 1: a grid_rows x (N*per_rank/grid_rows) mosaic instead of a strip (per_rank must be a multiple of grid_rows)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from imagestitch_b200 import sharded, stitching as S, synth

rows, cols, per_rank = (int(v) for v in (sys.argv[1:4] + ["800", "1200", "3"])[:3])
grid_rows = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = per_rank * world
fw = max(1.2, 0.75 * (n // grid_rows) / (2 * 5.9) / 0.5)       # keep the strip below ~340 degrees
Ks, Rs, scale = synth.strip_cameras(n, cols, rows, fw, 0.25, grid_rows=grid_rows)
be = sharded.GpuBackend(local)
st0 = S.Stitcher(be.ctx, "cylindrical", "dp", 5, S.WEIGHT_32F)
corners, sizes, roi = st0.plan([(cols, rows)] * n, Ks, Rs, scale)
plan = sharded.ShardPlan.build(corners, sizes, roi, world, 5)
mine = [synth.make_image(i, cols, rows, Ks[i], Rs[i], device=f"cuda:{local}") for i in range(n) if plan.owner[i] == rank]
sh = sharded.ShardedStitcher(be, sharded.Comm(dist), 5)
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    res = sh.stitch(mine, Ks, Rs, scale, plan)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
if rank == 0:
    print(f"sharded: world={world} images={n} {rows}x{cols} pano={roi[3]}x{roi[2]} cuts={plan.cuts} step={dt * 1e3:.2f} ms "
          f"speculation={sh.info['seam_speculation']} needed={sh.info['needed_images']}", flush=True)
# assemble on rank 0 and compare with the single-GPU pipeline
sums = [None] * world
dist.all_gather_object(sums, {i: int(t.to(torch.int64).sum().item()) for i, t in zip([k for k in range(n) if plan.owner[k] == rank], mine)})
parts = [None] * world
dist.all_gather_object(parts, (res["x0"], res["x1"], res["pano"].cpu().numpy(), res["pano_mask"].cpu().numpy(),
                               {i: m.cpu().numpy() for i, m in res["seam_masks"].items()}))
if rank == 0:
    parts.sort(key=lambda p: p[0])
    pano = np.concatenate([p[2] for p in parts], axis=1)
    pmask = np.concatenate([p[3] for p in parts], axis=1)
    imgs = [synth.make_image(i, cols, rows, Ks[i], Rs[i], device="cuda:0") for i in range(n)]
    for d_ in sums:
        for i, v in d_.items():
            v0 = int(imgs[i].to(torch.int64).sum().item())
            if v != v0:
                print(f"SOURCE image {i} differs between GPUs: byte sum {v} vs {v0}", flush=True)
    ref = st0.stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
    for p_ in parts:
        for i, m in p_[4].items():
            rmk = ref["seam_masks"][i].cpu().numpy()
            if not np.array_equal(m, rmk):
                d = np.argwhere(m != rmk)
                print(f"seam mask {i} differs: {len(d)} px, cols {d[:,1].min()}..{d[:,1].max()} rows {d[:,0].min()}..{d[:,0].max()}", flush=True)
    rp, rm = ref["pano"].cpu().numpy(), ref["pano_mask"].cpu().numpy()
    ok = np.array_equal(pano, rp) and np.array_equal(pmask, rm)
    print("SHARDED == SINGLE GPU:", ok, "single-GPU speculation:", be.ctx.seam_speculation, flush=True)
    if not ok:
        bad = np.argwhere((pano != rp).any(axis=2))
        print("mismatching pixels:", len(bad), "columns", bad[:, 1].min(), "..", bad[:, 1].max(), "rows", bad[:, 0].min(), "..", bad[:, 0].max(),
              "mask mismatches:", int((pmask != rm).sum()), flush=True)
    if not ok:
        sys.exit(1)
dist.barrier()
dist.destroy_process_group()
