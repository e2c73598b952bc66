"""Column-strip sharding of one panorama across the GPUs of a box.

One process per GPU (torch.distributed: NCCL on GPUs, gloo in the CPU tests).  Rank r owns the images whose centres
fall into its group of panorama columns (a contiguous run of a strip's images; whole columns of a 2-D mosaic) and the
panorama columns [cut[r], cut[r+1]).  Per step:

  warp        local (each rank warps its own images)
  X1          the right image of every pair that straddles a strip boundary travels to the owner of the left image
              (warped u8x3 + warped mask) -- P2P between column neighbours
  seam        every rank runs the pairs it owns CONCURRENTLY and speculatively on the entry masks
              (is_seam_pair_run); results needed elsewhere travel (X2: mask-sized messages); each pair with an
              earlier neighbour in the reference's order proves its result on the masks it would really have seen
              (is_seam_pair_check); one all-reduce of the verdict; if any proof fails every rank falls back to the
              reference's sequential loop on all-gathered inputs
  X3          blend halo: images that reach into a neighbour's strip travel with their final masks
  blend       is_blender_blend_strip on the rank's columns; the panorama stays sharded

The host logic here (ownership, exchange schedule, strip cuts) is pure Python over a small `backend` object, so the
N > 1 path is covered on CPU with gloo (tests/test_sharded_gloo.py) while the GPU backend calls the C ABI.
"""
from __future__ import annotations

import os
import threading
from dataclasses import dataclass, field

import numpy as np


def _overlap(c1, s1, c2, s2):
    return max(c1[0], c2[0]) < min(c1[0] + s1[0], c2[0] + s2[0]) and max(c1[1], c2[1]) < min(c1[1] + s1[1], c2[1] + s2[1])


@dataclass
class ShardPlan:
    """Everything every rank can derive from the geometry alone (deterministic, identical on all ranks)."""
    corners: list
    sizes: list
    roi: tuple
    world: int
    num_bands: int
    owner: list = field(default_factory=list)
    pairs: list = field(default_factory=list)        # active pairs in the reference's order ([SEAM]:100-111)
    cuts: list = field(default_factory=list)         # panorama columns (relative to roi.x): rank r owns [cuts[r], cuts[r+1])
    pairs_depend: bool = False                       # set (on every rank alike) once a step's proofs have failed: see stitch_general

    @staticmethod
    def build(corners, sizes, roi, world, num_bands):
        n = len(corners)
        assert n >= world, "fewer images than ranks"
        # n images over `world` ranks: the first n % world ranks take one image more
        counts = [n // world + (1 if r < n % world else 0) for r in range(world)]
        first = [sum(counts[:r]) for r in range(world + 1)]                  # rank r takes positions [first[r], first[r + 1]) of the x order
        p = ShardPlan([tuple(int(v) for v in c) for c in corners], [tuple(int(v) for v in s) for s in sizes], tuple(int(v) for v in roi), world, num_bands)
        # ownership by panorama column: images ordered by the x of their centre, n/world of them per rank.  For a strip
        # that is rank = i // m; for a 2-D mosaic (several rows of images) every rank gets the images of its columns.
        order = sorted(range(n), key=lambda i: (2 * p.corners[i][0] + p.sizes[i][0], i))
        p.owner = [0] * n
        for r in range(world):
            for k in range(first[r], first[r + 1]):
                p.owner[order[k]] = r
        allp = [(i, j) for i in range(n) for j in range(i + 1, n)][::-1]
        p.pairs = [(i, j) for (i, j) in allp if _overlap(p.corners[i], p.sizes[i], p.corners[j], p.sizes[j])]
        nb = min(num_bands, int(np.ceil(np.log(max(roi[2], roi[3])) / np.log(2.0))))
        q = 1 << nb
        cuts = [0]
        for r in range(1, world):
            lo = min(p.corners[i][0] for i in range(n) if p.owner[i] == r)                       # leftmost column of rank r
            hi = max(p.corners[i][0] + p.sizes[i][0] for i in range(n) if p.owner[i] == r - 1)   # rightmost of rank r-1
            mid = (lo + hi) // 2 - roi[0]
            mid = max(cuts[-1] + q, min(roi[2] - q, (mid // q) * q))
            cuts.append(mid)
        cuts.append(roi[2])
        p.cuts = cuts
        return p

    def pair_owner(self, k):
        return self.owner[self.pairs[k][0]]           # the rank that owns the left (lower-index) image

    def earlier(self, k, img):
        """indices of the earlier pairs (reference order) that contain image img"""
        return [q for q in range(k) if img in self.pairs[q]]


class Comm:
    """Neighbour exchange over torch.distributed (batch_isend_irecv); world == 1 degenerates to nothing."""

    def __init__(self, dist=None):
        self.dist = dist
        self.rank = dist.get_rank() if dist else 0
        self.world = dist.get_world_size() if dist else 1

    def exchange(self, sends, recvs):
        """sends: [(dst_rank, tensor)], recvs: [(src_rank, tensor)] -- matched by order per peer."""
        if not self.dist or (not sends and not recvs):
            return
        ops = [self.dist.P2POp(self.dist.isend, t, dst) for dst, t in sends] + [self.dist.P2POp(self.dist.irecv, t, src) for src, t in recvs]
        for req in self.dist.batch_isend_irecv(ops):
            req.wait()

    def post(self, sends, recvs):
        """start an exchange (same arguments as exchange); finish it with wait()"""
        if not self.dist or (not sends and not recvs):
            return []
        ops = [self.dist.P2POp(self.dist.isend, t, dst) for dst, t in sends] + [self.dist.P2POp(self.dist.irecv, t, src) for src, t in recvs]
        return self.dist.batch_isend_irecv(ops)

    def wait(self, pending):
        for req in pending:
            req.wait()

    def all_min(self, value: int, device):
        return self.all_min_result(self.all_min_start(value, device))

    def all_min_start(self, value: int, device):
        """the all-reduce (MIN) of a verdict, started now and read later with all_min_result(): nobody waits for the slowest rank
        in the middle of a step"""
        if not self.dist:
            return value
        import torch
        t = torch.tensor([value], dtype=torch.int32, device=device)
        work = self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, async_op=True)
        return (t, work)

    def all_min_result(self, pending):
        if not self.dist:
            return pending
        t, work = pending
        work.wait()
        return int(t.item())


class ShardedStitcher:
    """backend: object with warp / pair_run / pair_check / mask_and / clone / empty / strip_needs / blend_strip /
    run_concurrently / sync / device (see GpuBackend below and the oracle-backed double in the tests)."""

    def __init__(self, backend, comm: Comm, num_bands=5):
        self.be, self.comm, self.num_bands = backend, comm, num_bands
        self.info = {}

    def is_strip(self, plan: ShardPlan):
        """A left-to-right strip whose only cross-rank pairs join the last image of one rank to the first of the next."""
        n = len(plan.corners)
        if any(plan.owner[i] > plan.owner[i + 1] for i in range(n - 1)):       # ranks own consecutive index ranges, in order
            return False
        return all(j == i + 1 for (i, j) in plan.pairs)

    def stitch(self, my_images, Ks_all, Rs_all, scale, plan: ShardPlan):
        if hasattr(self.be, "seam_find_list") and self.is_strip(plan) and os.environ.get("IS_SHARD_GENERAL") != "1":
            return self.stitch_strip(my_images, Ks_all, Rs_all, scale, plan)
        return self.stitch_general(my_images, Ks_all, Rs_all, scale, plan)

    def stitch_strip(self, my_images, Ks_all, Rs_all, scale, plan: ShardPlan):
        """Strips: every rank hands its images plus the first image of its right neighbour to the batched seam finder in ONE call
        (the pair loop of the reference restricted to this rank: its own pairs and the boundary pair, proven equal to the
        sequential loop inside the library).  What the restriction misses is exactly one dependency per boundary -- in the
        reference's order the boundary pair (b, b+1) comes after (b+1, b+2), i.e. after the neighbour's own pairs have cleared
        part of image b+1 -- and that is checked afterwards with the mask the neighbour ends up with
        (is_seam_pair_same_structure).  Messages: X1 the neighbour's first image + mask, X2 its mask after its own pairs and,
        the other way, its mask after the boundary pair, one all-reduce of the verdict, X3 the blend halo."""
        be, comm, rank = self.be, self.comm, self.comm.rank
        import time
        dbg = os.environ.get("IS_SHARD_DEBUG") == "1"
        t_last = [time.perf_counter()]
        laps = {}

        def lap(name):
            if dbg:
                be.sync()
                t = time.perf_counter()
                laps[name] = laps.get(name, 0.0) + (t - t_last[0]) * 1e3
                t_last[0] = t
        n = len(plan.corners)
        mine = [i for i in range(n) if plan.owner[i] == rank]
        a, b = mine[0], mine[-1]
        warped, mask = {}, {}
        has_right = b + 1 < n and (b, b + 1) in plan.pairs
        has_left = a > 0 and (a - 1, a) in plan.pairs
        # ---- warp; X1 (my first image + entry mask -> left neighbour, for its boundary pair) is posted as soon as that image is
        #      warped, so that it travels while the other images are still being warped
        warped[a], mask[a] = be.warp(my_images[0], Ks_all[a], Rs_all[a], scale)
        sends, recvs = [], []
        if has_left:
            sends += [(rank - 1, warped[a]), (rank - 1, mask[a])]
        if has_right:
            warped[b + 1] = be.empty((plan.sizes[b + 1][1], plan.sizes[b + 1][0], 3), np.uint8)
            halo_entry = be.empty((plan.sizes[b + 1][1], plan.sizes[b + 1][0]), np.uint8)
            recvs += [(rank + 1, warped[b + 1]), (rank + 1, halo_entry)]
        x1_pending = comm.post(sends, recvs)
        for i, img in zip(mine[1:], my_images[1:]):
            warped[i], mask[i] = be.warp(img, Ks_all[i], Rs_all[i], scale)
        lap("warp")
        x0, x1 = plan.cuts[rank], plan.cuts[rank + 1]
        needed_by = [[i for i in range(n) if be.strip_needs(plan.sizes[i], plan.corners[i], plan.roi, self.num_bands, plan.cuts[r], plan.cuts[r + 1])]
                     for r in range(comm.world)]
        bh = be.blend_begin(plan.roi, self.num_bands)
        for i in mine:                                # image pyramids on a side stream while the seam stage runs; the masks are read at blend time
            if i in needed_by[rank]:
                be.blend_feed_early(bh, warped[i], mask[i], plan.corners[i], i + 1)
        comm.wait(x1_pending)
        # the neighbour's first image is here: if my strip blends it, its pyramid is built during the seam stage as well; its mask
        # buffer (halo_true) is filled by X2 and completed by X3, in place, before the blend reads it
        halo_true, halo_fed = None, False
        if has_right:
            halo_true = be.empty((plan.sizes[b + 1][1], plan.sizes[b + 1][0]), np.uint8)
            if (b + 1) in needed_by[rank]:
                be.blend_feed_early(bh, warped[b + 1], halo_true, plan.corners[b + 1], b + 2)
                halo_fed = True
        lap("x1")
        # ---- seam: one batched call over my images + the halo image
        entry_b = be.copy_of(mask[b]) if has_right else None
        ids = mine + ([b + 1] if has_right else [])
        halo_mask = be.copy_of(halo_entry) if has_right else None
        masks_in = [mask[i] for i in mine] + ([halo_mask] if has_right else [])
        self.info["seam"] = be.seam_find_list([warped[i] for i in ids], [plan.corners[i] for i in ids], masks_in)
        lap("seam")
        # ---- X2: my first image's mask after my own pairs -> left neighbour (its check); the halo's mask after the boundary pair -> its owner
        sends, recvs = [], []
        if has_left:
            after_own = be.copy_of(mask[a])           # mask[a] receives the boundary pair's clears below
            from_left = be.empty((plan.sizes[a][1], plan.sizes[a][0]), np.uint8)
            sends.append((rank - 1, after_own))
            recvs.append((rank - 1, from_left))
        if has_right:
            sends.append((rank + 1, halo_mask))
            recvs.append((rank + 1, halo_true))
        comm.exchange(sends, recvs)
        lap("x2")
        ok = 1
        if has_right:                                 # the boundary pair on the mask image b+1 really has at that point of the reference's loop
            ok = 1 if be.pair_same_structure(entry_b, halo_entry, halo_true, plan.corners[b], plan.corners[b + 1]) else 0
        # The verdict of all ranks (MIN) is only READ at the end of the step: the halo exchange and the blend are queued on the
        # assumption that every boundary pair is proven, which is the rule; a rank that waited here would wait for the slowest
        # rank of the box in the middle of every step.
        verdict = comm.all_min_start(ok, be.device)
        lap("check")
        if has_left:                                  # the boundary pair's clears (computed by the left neighbour) inside its rectangle
            i, j = a - 1, a
            rx0 = max(plan.corners[i][0], plan.corners[j][0]) - plan.corners[j][0]
            ry0 = max(plan.corners[i][1], plan.corners[j][1]) - plan.corners[j][1]
            rx1 = min(plan.corners[i][0] + plan.sizes[i][0], plan.corners[j][0] + plan.sizes[j][0]) - plan.corners[j][0]
            ry1 = min(plan.corners[i][1] + plan.sizes[i][1], plan.corners[j][1] + plan.sizes[j][1]) - plan.corners[j][1]
            be.mask_and(mask[a], from_left, rect=(rx0, ry0, rx1, ry1))
        final = {i: mask[i] for i in mine}
        if has_right:
            final[b + 1] = halo_true                  # completed below by X3 where the blend needs it
        lap("final_masks")
        # ---- X3: blend halo (images that reach into a neighbour's strip, with their final masks)
        # a receiver that already holds the image tells nobody: the sender must skip it too -- decide from the geometry alone
        sends, recvs = self._strip_x3(plan, needed_by, mine, warped, final, rank, comm, be, has_right, b)
        comm.exchange(sends, recvs)
        lap("x3")
        for i in needed_by[rank]:
            if i not in mine and not (halo_fed and i == b + 1):
                be.blend_feed(bh, warped[i], final[i], plan.corners[i], i + 1)
        pano, pmask = be.blend_finish(bh, x0, x1)
        self.info["needed_images"] = needed_by[rank]
        lap("blend")
        ok = comm.all_min_result(verdict)
        if os.environ.get("IS_SHARDED_FORCE_FALLBACK") == "1":
            ok = 0
        self.info["seam_speculation"] = ok
        if not ok:                                    # rare by construction: redo the step through the general path
            return self.stitch_general(my_images, Ks_all, Rs_all, scale, plan)
        if dbg:
            self.info["laps_ms"] = laps
            print(f"[shard rank {rank}] " + " ".join(f"{k}={v:.2f}" for k, v in laps.items()), flush=True)
        return dict(pano=pano, pano_mask=pmask, x0=x0, x1=x1, seam_masks={i: final[i] for i in mine})

    def _strip_x3(self, plan, needed_by, mine, warped, final, rank, comm, be, has_right, b):
        """X3 message lists.  Rank r already holds image b_r + 1 (X1), so only that image's FINAL mask travels to it; every other
        halo image travels whole.  Both sides derive the same schedule from the plan."""
        n = len(plan.corners)
        sends, recvs = [], []
        for r in range(comm.world):
            b_r = max(i for i in range(n) if plan.owner[i] == r)              # the last image of rank r
            for i in needed_by[r]:
                src = plan.owner[i]
                if src == r:
                    continue
                holds = (i == b_r + 1) and ((b_r, b_r + 1) in plan.pairs)
                if rank == src:
                    if not holds:
                        sends.append((r, warped[i]))
                    sends.append((r, final[i]))
                if rank == r:
                    if not holds:
                        warped[i] = be.empty((plan.sizes[i][1], plan.sizes[i][0], 3), np.uint8)
                        recvs.append((src, warped[i]))
                    if not (holds and i in final):            # the held image's mask buffer is already known to the blender (early feed): fill it in place
                        final[i] = be.empty((plan.sizes[i][1], plan.sizes[i][0]), np.uint8)
                    recvs.append((src, final[i]))
        return sends, recvs

    def stitch_general(self, my_images, Ks_all, Rs_all, scale, plan: ShardPlan):
        be, comm, rank = self.be, self.comm, self.comm.rank
        import time
        dbg = os.environ.get("IS_SHARD_DEBUG") == "1"
        t_last = [time.perf_counter()]
        laps = {}

        def lap(name):               # wall clock per phase (the phases end with a device sync already)
            if dbg:
                be.sync()
                t = time.perf_counter()
                laps[name] = laps.get(name, 0.0) + (t - t_last[0]) * 1e3
                t_last[0] = t
        n = len(plan.corners)
        mine = [i for i in range(n) if plan.owner[i] == rank]
        assert len(mine) == len(my_images)
        # ---- warp (local)
        warped, mask0 = {}, {}
        for i, img in zip(mine, my_images):
            warped[i], mask0[i] = be.warp(img, Ks_all[i], Rs_all[i], scale)
        be.sync()
        lap("warp")
        # which images every strip needs is geometry only; with a backend that can take feeds early, this rank's own images go
        # to the blender now (image pyramids on a side stream while the seam stage runs) with their final masks still to come
        x0, x1 = plan.cuts[rank], plan.cuts[rank + 1]
        needed_by = [[i for i in range(n) if be.strip_needs(plan.sizes[i], plan.corners[i], plan.roi, self.num_bands, plan.cuts[r], plan.cuts[r + 1])]
                     for r in range(comm.world)]
        early = hasattr(be, "blend_begin")
        final_buf, bh = {}, None
        if early:
            bh = be.blend_begin(plan.roi, self.num_bands)
            for i in mine:
                final_buf[i] = be.empty((plan.sizes[i][1], plan.sizes[i][0]), np.uint8)
                if i in needed_by[rank]:
                    be.blend_feed_early(bh, warped[i], final_buf[i], plan.corners[i], i + 1)
        # A panorama geometry whose pairs were found to depend on each other (the overlaps of a mosaic meet at the grid corners) goes
        # straight to the reference's loop on the following steps: the speculative runs and their proofs would fail again.
        speculate = not getattr(plan, "pairs_depend", False) or os.environ.get("IS_SHARD_ALWAYS_SPECULATE") == "1"
        # ---- X1: right images of boundary pairs -> owner of the left image
        sends, recvs = [], []
        seen = set()
        for k, (i, j) in enumerate(plan.pairs if speculate else []):
            src, dst = plan.owner[j], plan.pair_owner(k)
            if src == dst or (j, dst) in seen:
                continue
            seen.add((j, dst))
            if rank == src:
                sends += [(dst, warped[j]), (dst, mask0[j])]
            if rank == dst:
                warped[j] = be.empty((plan.sizes[j][1], plan.sizes[j][0], 3), np.uint8)
                mask0[j] = be.empty((plan.sizes[j][1], plan.sizes[j][0]), np.uint8)
                recvs += [(src, warped[j]), (src, mask0[j])]
        comm.exchange(sends, recvs)
        lap("x1")
        # ---- seam: speculative runs of the pairs this rank owns
        my_pairs = [k for k in range(len(plan.pairs)) if plan.pair_owner(k) == rank] if speculate else []
        outs, handles = {}, {}

        def run(k):
            i, j = plan.pairs[k]
            oi, oj, h = be.pair_run(warped[i], warped[j], plan.corners[i], plan.corners[j], mask0[i], mask0[j])
            outs[(k, i)], outs[(k, j)], handles[k] = oi, oj, h

        be.run_concurrently(run, my_pairs)
        be.sync()
        lap("seam_runs")
        # ---- X2: pair results to whoever needs them (final masks of the images' owners, validation of later pairs)
        sends, recvs = [], []
        for k, (i, j) in enumerate(plan.pairs if speculate else []):
            src = plan.pair_owner(k)
            for img in (i, j):
                dsts = {plan.owner[img]} | {plan.pair_owner(k2) for k2 in range(k + 1, len(plan.pairs)) if img in plan.pairs[k2]}
                for dst in sorted(dsts - {src}):
                    if rank == src:
                        sends.append((dst, outs[(k, img)]))
                    if rank == dst:
                        outs[(k, img)] = be.empty((plan.sizes[img][1], plan.sizes[img][0]), np.uint8)
                        recvs.append((src, outs[(k, img)]))
        comm.exchange(sends, recvs)
        lap("x2")
        # ---- validation of the owned pairs that have an earlier neighbour
        verdict = {}

        def check(k):
            i, j = plan.pairs[k]
            true = {}
            for img in (i, j):
                ev = plan.earlier(k, img)
                if not ev:
                    true[img] = mask0[img]
                    continue
                t = be.clone(outs[(ev[0], img)])
                for q in ev[1:]:
                    be.mask_and(t, outs[(q, img)])
                true[img] = t
            if true[i] is mask0[i] and true[j] is mask0[j]:
                verdict[k] = True
                return
            verdict[k] = be.pair_check(warped[i], warped[j], plan.corners[i], plan.corners[j], true[i], true[j], handles[k])

        be.run_concurrently(check, my_pairs)
        be.sync()
        lap("seam_checks")
        ok = comm.all_min(1 if all(verdict.values()) else 0, be.device) if speculate else 0
        if os.environ.get("IS_SHARDED_FORCE_FALLBACK") == "1":      # test hook: exercise the sequential fallback
            ok = 0
        if speculate and not ok and os.environ.get("IS_SHARDED_FORCE_FALLBACK") != "1":
            plan.pairs_depend = True
        self.info["seam_speculation"] = ok
        for k in my_pairs:
            be.pair_free(handles[k])
        final = {}
        if ok:
            for i in mine:                            # final mask = entry mask minus the clears of every pair it is in
                if early:
                    t = final_buf[i]
                    be.copy_into(t, mask0[i])
                else:
                    t = be.clone(mask0[i])
                for k, pr in enumerate(plan.pairs):
                    if i in pr:                       # a pair only clears pixels inside the intersection of its two images
                        a, b2 = pr
                        rx0 = max(plan.corners[a][0], plan.corners[b2][0]) - plan.corners[i][0]
                        ry0 = max(plan.corners[a][1], plan.corners[b2][1]) - plan.corners[i][1]
                        rx1 = min(plan.corners[a][0] + plan.sizes[a][0], plan.corners[b2][0] + plan.sizes[b2][0]) - plan.corners[i][0]
                        ry1 = min(plan.corners[a][1] + plan.sizes[a][1], plan.corners[b2][1] + plan.sizes[b2][1]) - plan.corners[i][1]
                        be.mask_and(t, outs[(k, i)], rect=(rx0, ry0, rx1, ry1))
                final[i] = t
        else:
            final, warped_all = self._sequential_fallback(plan, mine, warped, mask0)
            if early:                                 # the blender already holds this rank's images and mask buffers
                for i in mine:
                    be.copy_into(final_buf[i], final[i])
                    final[i] = final_buf[i]
                for i in range(n):
                    if i not in mine:
                        warped[i] = warped_all[i]
            else:
                warped = warped_all
        be.sync()
        lap("final_masks")
        # ---- X3: blend halo
        sends, recvs = [], []
        for r in range(comm.world):
            for i in needed_by[r]:
                src = plan.owner[i]
                if src == r:
                    continue
                if rank == src:
                    sends += [(r, warped[i]), (r, final[i])]
                if rank == r:
                    warped[i] = be.empty((plan.sizes[i][1], plan.sizes[i][0], 3), np.uint8)
                    final[i] = be.empty((plan.sizes[i][1], plan.sizes[i][0]), np.uint8)
                    recvs += [(src, warped[i]), (src, final[i])]
        comm.exchange(sends, recvs)
        lap("x3")
        # ---- blend my strip (feed order = global image order)
        if early:
            for i in needed_by[rank]:
                if i not in mine:
                    be.blend_feed(bh, warped[i], final[i], plan.corners[i], i + 1)
            pano, pmask = be.blend_finish(bh, x0, x1)
        else:
            feed = [(warped[i], final[i], plan.corners[i]) for i in needed_by[rank]]
            pano, pmask = be.blend_strip(feed, plan.roi, self.num_bands, x0, x1)
        self.info["needed_images"] = needed_by[rank]
        lap("blend")
        if dbg:
            self.info["laps_ms"] = laps
            print(f"[shard rank {rank}] " + " ".join(f"{k}={v:.2f}" for k, v in laps.items()), flush=True)
        return dict(pano=pano, pano_mask=pmask, x0=x0, x1=x1, seam_masks={i: final[i] for i in mine})

    def _sequential_fallback(self, plan, mine, warped, mask0):
        """A proof failed somewhere: all-gather the warped images + entry masks and run the reference's loop on every
        rank (redundantly); rare by construction (the pairs of a panorama are independent in practice)."""
        be, comm = self.be, self.comm
        n = len(plan.corners)
        allw, allm = {}, {}
        for i in range(n):
            src = plan.owner[i]
            sends, recvs = [], []
            if comm.rank == src:
                allw[i], allm[i] = warped[i], mask0[i]
                sends = [(r, t) for r in range(comm.world) if r != src for t in (warped[i], mask0[i])]
            else:
                allw[i] = be.empty((plan.sizes[i][1], plan.sizes[i][0], 3), np.uint8)
                allm[i] = be.empty((plan.sizes[i][1], plan.sizes[i][0]), np.uint8)
                recvs = [(src, allw[i]), (src, allm[i])]
            comm.exchange(sends, recvs)
        masks = be.seam_find_all([allw[i] for i in range(n)], plan.corners, [be.clone(allm[i]) for i in range(n)])
        return {i: masks[i] for i in range(n)}, allw


class GpuBackend:
    """The C ABI on one GPU; arrays are torch CUDA tensors."""

    def __init__(self, device_index, projection="cylindrical", weight_type=None, workers=6, ctx=None):
        import torch

        from . import stitching as S
        self.torch, self.S = torch, S
        self.device = torch.device("cuda", device_index)
        self.proj = projection
        self.wt = S.WEIGHT_32F if weight_type is None else weight_type
        self.ctx = ctx if ctx is not None else S.Context(device_index, use_torch_stream=True)
        self._nworkers = workers
        self._workers = None                  # contexts of the general path's worker threads, created on first use
        self.tls = threading.local()

    @property
    def workers(self):
        if self._workers is None:
            self._workers = [self.S.Context(self.device.index) for _ in range(self._nworkers)]
        return self._workers

    def seam_find_list(self, images, corners, masks):
        """the batched pair loop over a list of device images; masks are updated in place -> {path, waves}"""
        self.S.DpSeamFinder(self.ctx, "COLOR").find(images, corners, masks)      # device masks are updated in place
        return {"path": self.ctx.seam_path, "waves": self.ctx.seam_waves}

    def copy_of(self, t):
        return t.clone()                      # same stream as everything else of the strip path: no synchronisation

    def pair_same_structure(self, mask_i, mask_j_a, mask_j_b, tl_i, tl_j):
        return self.S.DpSeamFinder(self.ctx, "COLOR").pair_same_structure(mask_i, mask_j_a, mask_j_b, tl_i, tl_j)

    def _ctx(self):
        return getattr(self.tls, "ctx", self.ctx)

    def sync(self):
        self.torch.cuda.synchronize(self.device)

    def empty(self, shape, dtype):
        return self.torch.empty(shape, dtype={np.uint8: self.torch.uint8, np.int16: self.torch.int16}[dtype], device=self.device)

    def clone(self, t):
        r = t.clone()
        # the copy runs on torch's stream of the calling thread; the library may touch it next on another stream
        self.torch.cuda.current_stream(self.device).synchronize()
        return r

    def warp(self, img, K, R, scale):
        _, w, m = self.S.RotationWarper(self.ctx, self.proj, scale).warp_with_mask(img, K, R)
        return w, m

    def pair_run(self, wi, wj, ci, cj, mi, mj):
        return self.S.DpSeamFinder(self._ctx(), "COLOR").pair_run(wi, wj, ci, cj, mi, mj)

    def pair_check(self, wi, wj, ci, cj, mi, mj, h):
        return self.S.DpSeamFinder(self._ctx(), "COLOR").pair_check(wi, wj, ci, cj, mi, mj, h)

    def pair_free(self, h):
        self.S.DpSeamFinder(self.ctx, "COLOR").pair_free(h)

    def mask_and(self, dst, src, rect=None):
        if rect is not None:      # dst[src == 0] = 0 inside rect = (x0, y0, x1, y1): one small torch op on the main stream
            x0, y0, x1, y1 = rect
            if x1 > x0 and y1 > y0:
                dst[y0:y1, x0:x1].masked_fill_(src[y0:y1, x0:x1] == 0, 0)
            return
        c = self._ctx()
        self.S.DpSeamFinder(c, "COLOR").mask_and(dst, src)
        if c is not self.ctx:
            c.synchronize()

    def seam_find_all(self, images, corners, masks):
        return self.S.DpSeamFinder(self.ctx, "COLOR").find(images, corners, masks)

    def run_concurrently(self, fn, items):
        """fn(item) on worker threads, each bound to its own context / CUDA stream (ctypes releases the GIL)."""
        self.sync()
        if len(items) <= 1:
            for it in items:
                fn(it)
            return
        errs = []

        def work(t):
            self.tls.ctx = self.workers[t]
            try:
                for it in items[t::len(self.workers)]:
                    fn(it)
                self.workers[t].synchronize()
            except Exception as e:      # noqa: BLE001 - re-raised below
                errs.append(e)

        th = [threading.Thread(target=work, args=(t,)) for t in range(min(len(self.workers), len(items)))]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if errs:
            raise errs[0]

    def _blender(self, roi, num_bands):
        b = self.S.MultiBandBlender(self.ctx, 0, num_bands, self.wt)
        b.prepare(roi)
        return b

    def strip_needs(self, size_wh, corner, roi, num_bands, x0, x1):
        key = (tuple(roi), num_bands)
        if getattr(self, "_needs_key", None) != key:
            self._needs_b, self._needs_key = self._blender(roi, num_bands), key
        return self._needs_b.strip_needs(size_wh, corner, x0, x1)

    def blend_strip(self, feed, roi, num_bands, x0, x1):
        b = self._blender(roi, num_bands)
        for (img, mask, corner) in feed:
            b.feed(img, mask, corner, borrow=True)
        return b.blend_strip(x0, x1)

    # feeds spread over the step: own images early (image pyramid on a side stream, mask buffer filled in later), halo images late
    def blend_begin(self, roi, num_bands):
        return self._blender(roi, num_bands)

    def blend_feed_early(self, b, img, mask_buffer, corner, key):
        b.feed(img, mask_buffer, corner, defer=True, key=key)

    def blend_feed(self, b, img, mask, corner, key):
        b.feed(img, mask, corner, borrow=True, key=key)

    def blend_finish(self, b, x0, x1):
        return b.blend_strip(x0, x1)

    def copy_into(self, dst, src):
        dst.copy_(src)          # torch's current stream = the stream the library's main context adopted
