"""cuFFT D2Z + Z2D pair time for the 7-smooth lengths just above the E13 window (GPU box only): which lengthened window is cheapest?"""
import sys, torch
need = int(sys.argv[1]) if len(sys.argv) > 1 else 1341662
cands = []
v = need
while len(cands) < 400 and v < need * 1.08:
    r = v
    for f in (2, 3, 5, 7):
        while r % f == 0: r //= f
    if r == 1: cands.append(v)
    v += 1
res = []
for n in cands:
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    for _ in range(3): y = torch.fft.irfft(torch.fft.rfft(x), n=n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): y = torch.fft.irfft(torch.fft.rfft(x), n=n)
    e1.record(); torch.cuda.synchronize()
    res.append((e0.elapsed_time(e1) / 20 * 1e3, n))
res.sort()
for us, n in res[:12]: print("n %8d (+%.2f%%)  rfft+irfft %.1f us" % (n, 100.0 * (n - need) / need, us))
print("first candidate: n %d  %.1f us" % (cands[0], [u for u, n in res if n == cands[0]][0]))
