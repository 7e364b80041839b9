import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N = 257**3; k = 16
Bh = torch.zeros((k, N), dtype=torch.complex128).pin_memory()
Bd = torch.empty((k, N), dtype=torch.complex128, device="cuda")
Bh[0, 5] = 1.0
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); Bd.copy_(Bh, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("H2D pinned %.3f s  %.1f GB/s" % (dt, Bh.numel() * 16 / dt / 1e9))
    torch.cuda.synchronize(); t = time.perf_counter(); Bh.copy_(Bd, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("D2H pinned %.3f s  %.1f GB/s" % (dt, Bh.numel() * 16 / dt / 1e9))
Bn = Bh.numpy().T
t = time.perf_counter(); a = np.any(Bn); print("np.any %.3f s" % (time.perf_counter() - t), a)
t = time.perf_counter(); a = Bn.any(axis=0); print("np.any axis0 %.3f s" % (time.perf_counter() - t))
t = time.perf_counter(); a = float(abs(Bn[:, 0]).max()); print("abs max col %.3f s" % (time.perf_counter() - t))
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); del x
t = time.perf_counter(); y = torch.empty(4345495808, dtype=torch.uint8, device="cuda"); torch.cuda.synchronize(); print("torch alloc 4.3GB %.4f s" % (time.perf_counter() - t))
