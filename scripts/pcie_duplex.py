"""Host link: H2D and D2H alone and at the same time (pinned buffers, two streams).  python scripts/pcie_duplex.py"""
import torch

dev = "cuda:0"
up_h = torch.empty(432_000_000, dtype=torch.uint8, pin_memory=True)
up_d = torch.empty_like(up_h, device=dev)
dn_d = torch.empty(762_000_000, dtype=torch.uint8, device=dev)
dn_h = torch.empty(762_000_000, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        up_d.copy_(up_h, non_blocking=True)
    with torch.cuda.stream(s2):
        dn_h.copy_(dn_d, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


for _ in range(2):
    t_up = timed(lambda: up_d.copy_(up_h, non_blocking=True))
    t_dn = timed(lambda: dn_h.copy_(dn_d, non_blocking=True))
    t_both = timed(both)
    print(f"H2D 432 MB {t_up:.2f} ms ({0.432 / t_up * 1e3:.1f} GB/s)  D2H 762 MB {t_dn:.2f} ms ({0.762 / t_dn * 1e3:.1f} GB/s)  both at once {t_both:.2f} ms "
          f"(sum alone {t_up + t_dn:.2f}, max alone {max(t_up, t_dn):.2f})", flush=True)
