"""Write-only / read-only / copy HBM bandwidth of this GPU with plain torch ops (context for roofline.frac_dram: the tape is a
pure write stream in the forward and a pure read stream in the adjoint, MEASURED_PEAKS.json holds the copy figure)."""
import torch
n = 1 << 30                                  # 4 GiB of float32
a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
def tm(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
t = tm(lambda: a.fill_(1.0)); print("write-only (fill_):   %.0f GB/s" % (4 * n / t / 1e9))
t = tm(lambda: a.zero_()); print("write-only (memset):  %.0f GB/s" % (4 * n / t / 1e9))
t = tm(lambda: a.sum()); print("read-only (sum):      %.0f GB/s" % (4 * n / t / 1e9))
t = tm(lambda: b.copy_(a)); print("copy (read + write):  %.0f GB/s" % (8 * n / t / 1e9))
