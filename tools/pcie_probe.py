"""PCIe probe (GPU box): pinned H2D / D2H bandwidth alone and concurrently -- the ceiling of bench.py's e2e."""
import time, torch
n_in, n_out = 1610612736, 1073741824
h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
def chunks():
    k = 8
    for i in range(k):
        with torch.cuda.stream(s1): d_in[i*n_in//k:(i+1)*n_in//k].copy_(h_in[i*n_in//k:(i+1)*n_in//k], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i*n_out//k:(i+1)*n_out//k].copy_(d_out[i*n_out//k:(i+1)*n_out//k], non_blocking=True)
t1, t2, t3, t4 = run(h2d), run(d2h), run(both), run(chunks)
print("H2D 1.61 GB: %.1f ms (%.1f GB/s)   D2H 1.07 GB: %.1f ms (%.1f GB/s)   both: %.1f ms   both in 8 chunks: %.1f ms" % (t1*1e3, n_in/t1/1e9, t2*1e3, n_out/t2/1e9, t3*1e3, t4*1e3))
print("=> e2e ceiling for 512 images/step: %.0f img/s (copies fully overlapped), %.0f img/s (copies serialised)" % (512/t3, 512/(t1+t2)))
