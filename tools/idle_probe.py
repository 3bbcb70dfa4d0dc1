"""How long does the first kernel after an idle gap take?  (event-timed tiny kernel after sleeps)"""
import sys, time
sys.path.insert(0, '/root/repo')
import torch
from aggforce_b200 import _engine, _lib
x = torch.zeros(1 << 20, device="cuda")
g = torch.zeros((97, 97), dtype=torch.float64, device="cuda")
for gap_us in (0, 50, 200, 1000, 5000):
    out = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        while (time.perf_counter() - t0) * 1e6 < gap_us:
            pass
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        _lib.call("agf_symmetrize", _engine.ptr(g), 97, _engine.stream_ptr())
        e1.record()
        _lib.call("agf_symmetrize", _engine.ptr(g), 97, _engine.stream_ptr())
        e2.record()
        torch.cuda.synchronize()
        out.append((e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2) * 1e3))
    a = sorted(o[0] for o in out)[10]; b = sorted(o[1] for o in out)[10]
    print(f"idle gap {gap_us:5d} us: first kernel {a:7.1f} us, second {b:7.1f} us (medians)")
