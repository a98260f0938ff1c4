# c2 A/B of the arena-ring depth + raw PCIe read-back bandwidth of this box (pinned, 8.29 MB and 256 MB copies)
mkdir -p gpurun_out
bash tools/gpu/r2_c2_ab.sh SGL_FEW_ARENAS=1
python - <<PY
import torch, time
for mb in (8.2944, 256):
    n = int(mb * 1e6)
    d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    k = 50 if mb < 100 else 8
    for _ in range(k): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("D2H %.1f MB: %.1f GB/s" % (mb, n * k / dt / 1e9))
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(k): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("H2D %.1f MB: %.1f GB/s" % (mb, n * k / dt / 1e9))
PY
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
