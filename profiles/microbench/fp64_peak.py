import torch, time
torch.backends.cuda.matmul.allow_tf32 = False
dev = 'cuda'
def bench(fn, flops, name, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name}: {best:.3f} ms  {flops/best*1e-9:.2f} TFLOP/s", flush=True)
for n in (2048, 4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    bench(lambda: a @ b, 2*n**3, f"cublas dgemm n={n}")
    az = torch.randn(n, n, dtype=torch.complex128, device=dev); bz = torch.randn(n, n, dtype=torch.complex128, device=dev)
    bench(lambda: az @ bz, 8*n**3, f"cublas zgemm n={n}")
# rank-k update shape like LU trailing update: (n x k) @ (k x n)
for k in (64, 128, 256):
    n = 4096
    az = torch.randn(n, k, dtype=torch.complex128, device=dev); bz = torch.randn(k, n, dtype=torch.complex128, device=dev)
    c = torch.randn(n, n, dtype=torch.complex128, device=dev)
    bench(lambda: torch.addmm(c, az, bz, alpha=-1, out=c), 8*n*n*k, f"cublas zgemm rank-{k} update n={n}")
# LU via cusolver for reference
for n in (1024, 4096):
    az = torch.randn(n, n, dtype=torch.complex128, device=dev)
    bench(lambda: torch.linalg.lu_factor(az), 8/3*n**3, f"cusolver zgetrf n={n}", reps=3)
    bz = torch.randn(8, n, n, dtype=torch.complex128, device=dev)
    bench(lambda: torch.linalg.lu_factor(bz), 8*8/3*n**3, f"torch batched(8) zgetrf n={n}", reps=3)
