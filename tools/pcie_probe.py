"""PCIe probe (GPU box): pinned H2D / D2H bandwidth alone and concurrently -- the ceiling of bench.py's e2e -- on one
GPU, and on N GPUs at once (host memory bandwidth / root-complex sharing: what bounds e2e at N > 1).
usage: tools/pcie_probe.py [N ...]      e.g. tools/pcie_probe.py 1 2 4 8"""
import sys, threading, time
import torch

n_in, n_out = 1610612736, 1073741824  # one bench step: 512 images of 512^2, fp64 pixels + int32 labels in, fp64 out


class Lane:
    def __init__(self, dev):
        self.dev = dev
        with torch.cuda.device(dev):
            self.h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True)
            self.h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True)
            self.d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda:%d" % dev)
            self.d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda:%d" % dev)
            self.s1, self.s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def h2d(self):
        with torch.cuda.stream(self.s1):
            self.d_in.copy_(self.h_in, non_blocking=True)

    def d2h(self):
        with torch.cuda.stream(self.s2):
            self.h_out.copy_(self.d_out, non_blocking=True)

    def both(self):
        self.h2d(); self.d2h()

    def sync(self):
        torch.cuda.synchronize(self.dev)


def run(lanes, what, reps=5):
    """seconds per repetition with every lane doing `what` at the same time (one host thread per GPU)"""
    bar = threading.Barrier(len(lanes) + 1)
    def work(l):
        torch.cuda.set_device(l.dev)
        getattr(l, what)(); l.sync()
        bar.wait()
        for _ in range(reps):
            getattr(l, what)()
        l.sync()
        bar.wait()
    ths = [threading.Thread(target=work, args=(l,)) for l in lanes]
    for t in ths: t.start()
    bar.wait(); t0 = time.perf_counter()
    bar.wait(); dt = time.perf_counter() - t0
    for t in ths: t.join()
    return dt / reps


counts = [int(a) for a in sys.argv[1:]] or [1]
have = torch.cuda.device_count()
lanes = [Lane(d) for d in range(min(max(counts), have))]
for n in counts:
    if n > have:
        print("N=%d: only %d GPUs visible" % (n, have)); continue
    ls = lanes[:n]
    t1, t2, t3 = run(ls, "h2d"), run(ls, "d2h"), run(ls, "both")
    print("N=%d  per GPU: H2D 1.61 GB %.1f ms (%.1f GB/s)  D2H 1.07 GB %.1f ms (%.1f GB/s)  both %.1f ms  => e2e ceiling (fp64 interface) "
          "%.0f images/s per GPU, %.0f in total; uint8 interface (0.40 + 0.13 GB) ~%.0f in total"
          % (n, t1 * 1e3, n_in / t1 / 1e9, t2 * 1e3, n_out / t2 / 1e9, t3 * 1e3, 512 / t3, n * 512 / t3, n * 512 / (t3 * 0.2)), flush=True)
