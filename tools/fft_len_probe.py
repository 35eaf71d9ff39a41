"""Which 7-smooth FFT lengths just above N does cuFFT run fastest (Z2D, batched)?  GPU box only."""
import sys, time, torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2677430
span = float(sys.argv[2]) if len(sys.argv) > 2 else 0.04
def smooth(lo, hi):
    out = []
    p2 = 1
    while p2 <= hi:
        p3 = p2
        while p3 <= hi:
            p5 = p3
            while p5 <= hi:
                p7 = p5
                while p7 <= hi:
                    if p7 >= lo and p7 % 2 == 0: out.append(p7)
                    p7 *= 7
                p5 *= 5
            p3 *= 3
        p2 *= 2
    return sorted(out)
cands = smooth(N, int(N * (1 + span)))
print(len(cands), "candidates")
batch = 128
res = []
for n in cands:
    x = torch.randn(batch, n // 2 + 1, dtype=torch.complex128, device="cuda")
    torch.fft.irfft(x, n=n); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): torch.fft.irfft(x, n=n)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    def fac(n):
        f = {}
        for p in (2, 3, 5, 7):
            while n % p == 0: f[p] = f.get(p, 0) + 1; n //= p
        return f
    res.append((ms / batch * 1e3, n, fac(n)))
    del x
for r in sorted(res)[:12]: print("%.1f us/lc  n=%d (+%.2f%%) %s" % (r[0], r[1], 100 * (r[1] / N - 1), r[2]))
print("first (smallest) candidate:", [r for r in res if r[1] == cands[0]])
