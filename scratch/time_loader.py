"""device loader: in-memory samples page-locked in place vs staged through a pinned buffer; .npy files"""
import os, sys, time, tempfile, numpy as np, torch
sys.path.insert(0, os.getcwd())
import scenenet_b200 as sb
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
base = [np.concatenate([rng.uniform(0, 30, (60_000, 3)) + np.array([544850.0, 4634550.0, 160.0]), rng.integers(1, 16, (60_000, 1)).astype(np.float64)], 1) for _ in range(32)]
def run(loader, tag):
    for _ in loader: pass
    torch.cuda.synchronize(); t0 = time.perf_counter(); nb = 0
    for _ in range(2):
        for x, y in loader: nb += x.shape[0]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{tag}: {nb / dt:.0f} clouds/s ({nb} clouds, {dt * 1e3 / (nb / 32):.2f} ms per batch of 32)", flush=True)
run(sb.TS40KDeviceLoader(base * 8, batch_size=32, device=dev), "in-memory, page-locked in place")
run(sb.TS40KDeviceLoader(base * 8, batch_size=32, device=dev, pin_sources_bytes=0), "in-memory, staged")
run(sb.TS40KDeviceLoader(base * 8, batch_size=32, device=dev, dtype=torch.int32), "in-memory, page-locked, packed bits out")
d = tempfile.mkdtemp()
paths = []
for i, a in enumerate(base):
    np.save(os.path.join(d, f"s{i}.npy"), a); paths.append(os.path.join(d, f"s{i}.npy"))
run(sb.TS40KDeviceLoader(paths * 8, batch_size=32, device=dev), ".npy files (page cache) -> pinned")
