"""How the host link splits between concurrent upload and download when the copies are issued in pieces (one GPU).
Prints, per variant, the time until each direction has finished and the average rates while both were active."""
import time

import torch

UP, DOWN = 269_499_592, 400_815_088
hu = torch.empty(UP, dtype=torch.uint8).pin_memory()
hd = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
du = torch.empty(UP, dtype=torch.uint8, device="cuda")
dd = torch.empty(DOWN, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up_piece, down_piece):
    e0 = torch.cuda.Event(enable_timing=True)
    eu, ed = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    s1.wait_event(e0)
    s2.wait_event(e0)
    with torch.cuda.stream(s1):
        for o in range(0, UP, up_piece):
            du[o:o + up_piece].copy_(hu[o:o + up_piece], non_blocking=True)
        eu.record()
    with torch.cuda.stream(s2):
        for o in range(0, DOWN, down_piece):
            hd[o:o + down_piece].copy_(dd[o:o + down_piece], non_blocking=True)
        ed.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(eu), e0.elapsed_time(ed)


MB = 1 << 20
for name, up_p, dn_p in [("whole / whole", UP, DOWN), ("25 MB / 25 MB", 25 * MB, 25 * MB), ("25 MB / 12 MB", 25 * MB, 12 * MB),
                         ("4 MB / whole", 4 * MB, DOWN), ("1 MB / whole", MB, DOWN), ("256 KB / whole", MB // 4, DOWN),
                         ("whole / 4 MB", UP, 4 * MB), ("2 MB / 50 MB", 2 * MB, 50 * MB)]:
    run(up_p, dn_p)
    r = [run(up_p, dn_p) for _ in range(4)]
    tu = sorted(x[0] for x in r)[1]
    td = sorted(x[1] for x in r)[1]
    both = min(tu, td)
    print(f"{name:16s} upload done {tu:6.2f} ms  download done {td:6.2f} ms   "
          f"(rates while both active ~ up {UP / tu / 1e6 if tu <= td else float('nan'):5.1f} GB/s, "
          f"down {DOWN / td / 1e6 if td <= tu else float('nan'):5.1f} GB/s)")
